# two-GPU session (gpurun --gpus 2 -- 'bash tools/gpu_session_two.sh'): the tests that need two devices, the tutorial programs, and the driver's launch line at N = 2
mkdir -p gpurun_out
nvidia-smi -L
( time python -m pytest tests/test_sharded_solver.py tests/test_tutorials_run_gpu.py tests/test_orchestration_gpu.py tests/test_coupled_gpu.py -m gpu -q -s ) > gpurun_out/r2f2_pytest.txt 2>&1; tail -8 gpurun_out/r2f2_pytest.txt; grep "tutorial" gpurun_out/r2f2_pytest.txt | head
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2f2_bench_n2.json 2> gpurun_out/r2f2_bench_n2.err; tail -5 gpurun_out/r2f2_bench_n2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2f2_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("N=2 HEADLINE value %.4e kernel_ms %.4f e2e %.4e (%.3f ms/step) copies-alone %.1f GB/s/rank binding %s" % (d["value"], r["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("copies_alone_gbs_per_rank", 0), d.get("host_binding")))
    for k,c in d.get("configs",{}).items():
        if "error" in c: print(k, c); continue
        print(k, "value %.4e kernel_ms %.4f e2e %.4e scaling %s" % (c["value"], c["roofline"]["kernel_ms"], c["e2e"]["value"], c.get("scaling")))
except Exception as e: print("bench unreadable", e)
PY
