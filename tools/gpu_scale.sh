# scaling runs of the default bench command (the driver's launch line) at N = 1, 2, 4, 8 on one box
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/scale_lorenz_n$n.json 2> gpurun_out/scale_lorenz_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_lorenz_n$n.json 2> gpurun_out/scale_lorenz_n$n.err; fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 5 --warmup 3 --workload coupled > gpurun_out/scale_coupled_n8.json 2> gpurun_out/scale_coupled_n8.err
python - <<PY
import json
base=None
for f in ("lorenz_n1","lorenz_n2","lorenz_n4","lorenz_n8","coupled_n8"):
    try:
        d=json.loads(open(f"gpurun_out/scale_{f}.json").read().strip().splitlines()[-1])
        if f=="lorenz_n1": base=d["value"]
        print(f, "n_gpus", d["n_gpus"], "value %.4e e2e %.4e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), ("eff %.3f" % (d["value"]/base/d["n_gpus"])) if base and f.startswith("lorenz") else "", d["clocks"])
    except Exception as e:
        print(f, "FAILED", e, open(f"gpurun_out/scale_{f}.err").read()[-800:])
PY
