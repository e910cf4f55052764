"""Per-source-line roll-up of an ncu report (needs -lineinfo and --import-source on): executed warp-instructions and
stall samples per file:line, top N.   usage: python tools/ncu_lines.py rep.ncu-rep [N] [warp_steps]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wsteps = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
path, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); continue
    if hdr and r[0].strip().isdigit():
        try: lines.append((path, int(r[0]), r[1].strip(), int(r[iS] or 0), int(r[iE] or 0)))
        except ValueError: pass
tot_s = sum(l[3] for l in lines) or 1; tot_e = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s}  total warp-instructions {tot_e}" + (f"  ({tot_e / wsteps:.1f} per warp-step)" if wsteps else ""))
byfile = {}
for l in lines:
    a = byfile.setdefault(l[0], [0, 0]); a[0] += l[3]; a[1] += l[4]
for f, (sm, ex) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  file {f:44s} samples {100.0 * sm / tot_s:5.1f}%  instructions {100.0 * ex / tot_e:5.1f}%")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{l[0][:28]:28s}:{l[1]:5d} samples {100.0 * l[3] / tot_s:5.1f}% inst {100.0 * l[4] / tot_e:5.1f}%" + (f" {l[4] / wsteps:7.1f}/step" if wsteps else "") + f"  | {l[2][:110]}")
