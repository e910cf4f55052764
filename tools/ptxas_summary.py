"""Registers, stack and spills of every integration kernel of the shipped build, from the ptxas logs the csrc Makefile writes
(-Xptxas -v).  usage: python tools/ptxas_summary.py > profiles/rN_ptxas_summary.txt   (no GPU needed)"""
import glob, os, re, subprocess
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "massively-parallel-gpu-ode-solver_b200", "csrc")
print("# ptxas -v of the shipped build (libmpgos_b200.so): registers | stack frame | spill stores | spill loads | static shared memory | kernel")
for log in sorted(glob.glob(os.path.join(CSRC, "builtin_*.ptxas.log"))):
    if ".strict." in log:
        continue
    txt = open(log).read()
    pat = (r"Compiling entry function '(\S+)' for 'sm_100a'\s*\nptxas info\s*: Function properties for \S+\s*\n\s*(\d+) bytes stack frame, "
           r"(\d+) bytes spill stores, (\d+) bytes spill loads\s*\nptxas info\s*: Used (\d+) registers(.*)")
    for m in re.finditer(pat, txt):
        name, stack, st, ld, regs, rest = m.groups()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if "perthread_kernel" not in dem and "perblock_kernel" not in dem:
            continue
        dem = re.sub(r"\(mpgos_b200::Per\w+Args.*$", "", dem).replace("mpgos_b200::", "").replace("void ", "")
        smem = re.search(r"(\d+) bytes smem", rest)
        print(f"{regs:>4} regs {stack:>4} B stack {st:>4} B st {ld:>4} B ld {smem.group(1) if smem else 0:>6} B smem | {dem}")
