# single-GPU session at the end of a round (gpurun -- 'bash tools/gpu_session_single.sh'): smoke(), the whole GPU suite, the bench line of both
# arms, the ncu launch list of the bench command, one --set full capture per configuration's kernel, compute-sanitizer over the round-2 cases.
# Outputs under gpurun_out/ (r2f_*); the summaries committed under profiles/ come from them.
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -6
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2f_pytest.txt 2>&1; tail -5 gpurun_out/r2f_pytest.txt
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2f_bench_reference_n1.json 2> gpurun_out/r2f_bench_reference.err; tail -c 600 gpurun_out/r2f_bench_reference_n1.json
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2f_bench_n1.json").read().strip().splitlines()[-1])
print("HEADLINE", j["value"], j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "peak", j["roofline"]["peak"], "ref", j.get("reference_gpu", {}).get("speedup_kernel"))
for k, c in j.get("configs", {}).items():
    print(k, "value %.4e" % c["value"], "kernel_ms", c.get("roofline", {}).get("kernel_ms"), "frac", c.get("roofline", {}).get("frac"), "vs ref", c.get("reference_gpu", {}).get("speedup_kernel"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu --configs none > gpurun_out/r2f_bench_under_ncu.log 2>&1
for w in lorenz impact kellermiksis_dense kellermiksis_dense_fm; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/r2f_prof_$w python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --configs none > /dev/null 2>&1
done
for w in coupled coupled_fm; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:perblock_kernel -s 6 -c 2 -f -o gpurun_out/r2f_prof_$w python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --configs none > /dev/null 2>&1
done
ls -la gpurun_out/r2f_prof_*.ncu-rep | cut -c25-
SANITIZE_CASES="round2" timeout 1500 bash tools/gpu_sanitize.sh 2>&1 | tail -12
