# full GPU check: parity tests, smoke, bench line (both arms), launch list and one full ncu capture
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference_r1.json 2>&1; cat gpurun_out/bench_reference_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/prof_lorenz_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
