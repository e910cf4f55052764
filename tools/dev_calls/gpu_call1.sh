set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
./build/microbench_dmma > gpurun_out/dmma.txt 2>&1; cat gpurun_out/dmma.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; cat gpurun_out/bench_s3.json
python tools/probe_refgpu.py > gpurun_out/probe_refgpu.txt 2>&1; tail -30 gpurun_out/probe_refgpu.txt
python tools/time_coupled.py 8192 2>&1 | tail -2
python tools/probe_coupled.py > gpurun_out/probe_coupled.txt 2>&1; tail -12 gpurun_out/probe_coupled.txt
ncu --set full --clock-control none --import-source on -k regex:perblock_kernel -c 1 -f -o gpurun_out/prof_coupled_s3 python tools/time_coupled.py 2048 > /dev/null 2>&1
ls -la gpurun_out
