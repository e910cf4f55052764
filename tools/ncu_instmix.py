"""Dynamic instruction mix of the profiled kernel from an ncu report's source page:
executed warp-instructions per opcode, per attempted step if --steps is given.
usage: python tools/ncu_instmix.py rep.ncu-rep [warp_steps]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
wsteps = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, samples, total = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= iE: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if not m: continue
    op = m.group(2)
    n = int(r[iE] or 0)
    mix[op] += n; total += n
    samples[op] += int(r[iSm] or 0)
print("total warp instructions", total, "per warp-step" if wsteps else "", total / wsteps if wsteps else "")
for op, n in mix.most_common(30):
    print(f"{op:12s} {n:14d} {100.0*n/total:6.2f}%  " + (f"{n/wsteps:8.2f}/step" if wsteps else "") + f"  samples {samples[op]}")
