# ncu captures of every configuration's kernel (one launch each, --set full) + the launch list of the default bench command
TAG=${1:-s3}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_under_ncu.log 2>&1
for w in lorenz impact kellermiksis_dense; do
  ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$w python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:perblock_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_coupled python bench.py --workload coupled --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$TAG.csv
