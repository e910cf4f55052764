set -x
mkdir -p gpurun_out
python -m pytest tests/test_coupled_gpu.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_perthread_gpu.py -q -m gpu -k "bit_identical_to_reference" 2>&1 | grep -E "AssertionError|assert|passed|failed|^E" | cut -c1-300 | head -20
TC=0 python tools/time_coupled.py 8192 2>&1 | tail -1
TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
TC=1 python tools/time_coupled.py 65536 2>&1 | tail -1
python tools/probe_coupled.py 2>&1 | tail -4 | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:perblock_kernel -c 1 -f -o gpurun_out/prof_coupled_mma python tools/time_coupled.py 8192 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/prof_impact python bench.py --workload impact --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/prof_km python bench.py --workload kellermiksis --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
ls -la gpurun_out | tail -5
