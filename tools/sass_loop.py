"""Static view of a kernel's hot loop: histogram of SASS opcodes between the head of the widest
backward branch and the branch itself.  usage: sass_loop.py file.o 'substring of mangled or demangled name' [--dump]"""
import collections, re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
dump = "--dump" in sys.argv
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if key not in name and key not in dem:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\w+\s+)?)([A-Z0-9_]+)([^;]*);", f):
        ins.append((int(m.group(1), 16), m.group(3), (m.group(2) + m.group(3) + m.group(4)).strip()))
    best = None
    for a, op, text in ins:
        if op == "BRA":
            t = re.search(r"0x([0-9a-f]+)", text)
            if t and int(t.group(1), 16) < a:
                span = a - int(t.group(1), 16)
                if best is None or span > best[0]:
                    best = (span, int(t.group(1), 16), a)
    print("==", dem[:140])
    if not best:
        print("no loop"); continue
    _, lo, hi = best
    body = [(a, op, t) for a, op, t in ins if lo <= a <= hi]
    h = collections.Counter(op for _, op, _ in body)
    dp = sum(n for op, n in h.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "F2F", "DMNMX"))
    print(f"loop 0x{lo:x}-0x{hi:x}: {len(body)} instructions, FP64-pipe {dp}, other {len(body)-dp}")
    print("  " + "  ".join(f"{op}:{n}" for op, n in h.most_common()))
    if dump:
        for a, op, t in body:
            print(f"  {a:05x}  {t}")
