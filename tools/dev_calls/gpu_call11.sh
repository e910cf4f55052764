for i in 1 2 3; do python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   Assert|passed|failed" | cut -c1-200; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_s3_lorenz_n2.json 2> gpurun_out/bench_s3_lorenz_n2.err; tail -c 600 gpurun_out/bench_s3_lorenz_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload coupled > gpurun_out/bench_s3_coupled_n2.json 2> gpurun_out/bench_s3_coupled_n2.err; tail -c 600 gpurun_out/bench_s3_coupled_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 3 | cut -c1-200
python - <<PY
import json
for f in ("lorenz_n2","coupled_n2"):
    d=json.loads(open(f"gpurun_out/bench_s3_{f}.json").read().strip().splitlines()[-1])
    print(f, "n_gpus", d["n_gpus"], "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
PY
