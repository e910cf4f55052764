# 2 GPUs: the orchestration test that needs two devices, and the N=2 bench lines of the headline and the coupled configuration
python -m pytest tests/test_orchestration_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error|skipped" | cut -c1-300 | head
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/scale2_lorenz_n1.json 2> gpurun_out/scale2_lorenz_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale2_lorenz_n2.json 2> gpurun_out/scale2_lorenz_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --steps 5 --warmup 3 --workload coupled > gpurun_out/scale2_coupled_n2.json 2> gpurun_out/scale2_coupled_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29604 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/scale2_reference_n2.json 2> gpurun_out/scale2_reference_n2.err
for f in lorenz_n1 lorenz_n2 coupled_n2 reference_n2; do tail -1 gpurun_out/scale2_$f.json | cut -c1-400; done
