set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/prof_lorenz_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
