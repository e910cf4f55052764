# bench lines of every BASELINE configuration (product arm) + the reference arm of the headline one
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
TAG=${1:-s3}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_lorenz.json 2> gpurun_out/bench_${TAG}_lorenz.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_lorenz_reference.json 2>&1
python bench.py --workload lorenz --precision f32 --steps 10 > gpurun_out/bench_${TAG}_lorenz_f32.json 2> gpurun_out/bench_${TAG}_lorenz_f32.err
python bench.py --workload impact --phases 8 --steps 5 > gpurun_out/bench_${TAG}_impact.json 2> gpurun_out/bench_${TAG}_impact.err
python bench.py --workload kellermiksis_dense --phases 2 --steps 5 > gpurun_out/bench_${TAG}_kellermiksis_dense.json 2> gpurun_out/bench_${TAG}_kellermiksis_dense.err
python bench.py --workload duffing --phases 8 --steps 10 > gpurun_out/bench_${TAG}_duffing.json 2> gpurun_out/bench_${TAG}_duffing.err
python bench.py --workload coupled --steps 5 > gpurun_out/bench_${TAG}_coupled.json 2> gpurun_out/bench_${TAG}_coupled.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "UNREADABLE", e, open(f.replace('.json','.err')).read()[-600:] if not f.endswith('reference.json') else ''); continue
    r=d.get("roofline",{}); g=d.get("reference_gpu",{}); c=d.get("cpu_baseline",{})
    print(f.split("bench_")[1][:-5], "value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]), "kernel_ms %s frac %s" % (r.get("kernel_ms"), r.get("frac")), "refgpu x%s" % g.get("speedup_kernel"), "cpu %s (%s cores, %s)" % (c.get("value"), c.get("cores"), c.get("kind")), d.get("clocks"))
PY
