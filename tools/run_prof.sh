# usage: bash tools/run_prof.sh TAG [bench args...]   -- quick bench line + ncu full capture of one solver launch
TAG=$1; shift
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu "$@" > /dev/null 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("$TAG", "value %.3e"%d["value"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["clocks"])
PY
