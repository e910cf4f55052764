# eight-GPU session (gpurun --gpus 8 -- 'bash tools/gpu_session_eight.sh'): one solver object over 1 / 2 / 4 / 8 devices, and the driver's launch line at N = 8 (headline + C5 + strong scaling)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python tools/probe_sharded_scaling.py 2>&1 | tail -8 | tee gpurun_out/r2s8_sharded_object.txt

( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-reference-gpu --no-cpu-baseline --configs C5_coupled_bubbles_rkck45_f64,C5_coupled_bubbles_fastmath,C2_lorenz_rkck45_f64_strong_scaling ) > gpurun_out/r2s8_bench_n8.json 2> gpurun_out/r2s8_bench_n8.err; tail -4 gpurun_out/r2s8_bench_n8.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2s8_bench_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("N=8 HEADLINE value %.4e kernel_ms %.4f e2e %.4e (%.3f ms/step) copies-alone %.1f GB/s/rank binding %s" % (d["value"], r["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("copies_alone_gbs_per_rank", 0), d.get("host_binding")))
    for k,c in d.get("configs",{}).items():
        if "error" in c: print(k, c); continue
        print(k, "value %.4e kernel_ms %.4f e2e %.4e scaling %s" % (c["value"], c["roofline"]["kernel_ms"], c["e2e"]["value"], c.get("scaling")))
except Exception as e: print("bench unreadable", e)
PY
