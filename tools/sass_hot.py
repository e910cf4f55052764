"""Static issue-slot budget of a kernel's step loop, by source line (development tool).

On sm_100a the thread-per-system kernels are bound by instruction issue: an FP64-pipe instruction occupies the sub-partition's
dispatch port for two cycles, every other instruction for one (tools/microbench_fp64.cu), so a warp-step costs about
2 * N_fp64 + N_other cycles.  This script disassembles an object with `nvdisasm -g` (source line per instruction, the build has
-lineinfo), takes the instructions between the head of the widest backward branch and the branch (the step loop), drops the ones
that belong to source lines marked cold -- between `// [cold>` and `// <cold]` comments in the kernel source: the exact-arithmetic
fallbacks, device messages, limits -- and prints the budget per source line.

usage: sass_hot.py file.o 'substring of the mangled kernel name' [--lines] [--dump]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "F2F", "DMNMX", "MUFU.RSQ64H", "MUFU.RCP64H", "FRND", "I2F.F64", "F2I.F64", "DMMA")


def cold_lines(path):
    """lines between `[cold>` and `<cold]` (never on the ordinary step) or `[rare>` and `<rare]` (e.g. the event scans: on the
    steps around an event only)"""
    cold, on = set(), False
    try:
        for i, line in enumerate(open(path), 1):
            if "[cold>" in line or "[rare>" in line:
                on = True
            if on:
                cold.add(i)
            if "<cold]" in line or "<rare]" in line:
                on = False
    except OSError:
        pass
    return cold


def main():
    obj, key = sys.argv[1], sys.argv[2]
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
        text = "".join(subprocess.run(["nvdisasm", "-gi", "-c", c], capture_output=True, text=True).stdout for c in cubins)
    sections = re.split(r"\n//-+ \.text\.", text)
    cold_cache = {}
    for sec in sections[1:]:
        name = sec.split(" ", 1)[0]
        if key not in name:
            continue
        # with -gi every instruction is preceded by its inline chain: one "//## File .., line N [inlined at File .., line M]" per frame
        ins, chain, fresh = [], [(None, 0)], True
        for line in sec.splitlines():
            if line.lstrip().startswith("//## File"):
                frames = [(os.path.normpath(f), int(n)) for f, n in re.findall(r'"([^"]+)", line (\d+)', line)]
                if fresh:
                    chain, fresh = [], False
                chain.extend(frames)
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+((?:@!?U?P\w+\s+)?)([A-Z0-9_.]+)([^;]*);", line)
            if m:
                fresh = True
                ins.append((int(m.group(1), 16), m.group(3), (m.group(2) + m.group(3) + m.group(4)).strip(), tuple(chain)))
        labels = {}
        for line in sec.splitlines():
            m = re.match(r"\s*(\.L_x_\d+):", line)
            if m:
                labels[m.group(1)] = None
        # label addresses: the instruction following the label line
        addr_of, pending = {}, []
        for line in sec.splitlines():
            m = re.match(r"\s*(\.L_x_\d+):", line)
            if m:
                pending.append(m.group(1))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", line)
            if m and pending:
                for p in pending:
                    addr_of[p] = int(m.group(1), 16)
                pending = []
        best = None
        for a, op, t, _ in ins:
            if op.startswith("BRA"):
                m = re.search(r"(\.L_x_\d+)", t)
                if m and m.group(1) in addr_of and addr_of[m.group(1)] < a:
                    span = a - addr_of[m.group(1)]
                    if best is None or span > best[0]:
                        best = (span, addr_of[m.group(1)], a)
        print("==", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:150])
        if not best:
            print("no loop")
            continue
        _, lo, hi = best
        body = [x for x in ins if lo <= x[0] <= hi]
        per_line = collections.OrderedDict()
        tot = {"hot": [0, 0], "cold": [0, 0]}
        ops = collections.Counter()
        def is_cold(frames):
            for f, ln in frames:
                if f not in cold_cache:
                    cold_cache[f] = cold_lines(f) if f else set()
                if ln in cold_cache[f]:
                    return True
            return False

        for a, op, t, frames in body:
            f, ln = frames[0]
            kind = "cold" if is_cold(frames) else "hot"
            is64 = any(op == p or op.startswith(p + ".") for p in FP64)
            tot[kind][0 if is64 else 1] += 1
            if kind == "hot":
                ops[op.split(".")[0]] += 1
                k = (os.path.basename(f) if f else "?", ln)
                per_line.setdefault(k, [0, 0])[0 if is64 else 1] += 1
        h64, hot_other = tot["hot"]
        print(f"loop 0x{lo:x}-0x{hi:x}: {len(body)} instructions; hot lines: FP64-pipe {h64} + other {hot_other} -> {2 * h64 + hot_other} issue cycles per warp-step;"
              f" cold lines: {tot['cold'][0]} + {tot['cold'][1]}")
        print("  " + "  ".join(f"{op}:{n}" for op, n in ops.most_common()))
        if "--lines" in sys.argv:
            src = {}
            for (f, ln), (a, b) in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1])):
                print(f"  {f}:{ln:<5d} fp64 {a:3d} other {b:3d}  cycles {2 * a + b:4d}")
        if "--dump" in sys.argv:
            for a, op, t, frames in body:
                f, ln = frames[0]
                kind = "c" if is_cold(frames) else " "
                print(f"  {a:05x} {kind} {os.path.basename(f) if f else '?'}:{ln:<5d} {t}")


if __name__ == "__main__":
    main()
