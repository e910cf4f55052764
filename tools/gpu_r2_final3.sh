# round 2, last single-GPU session: smoke(), the whole GPU suite, both bench arms (what the driver runs at round end)
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -4
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2g_pytest.txt 2>&1; tail -4 gpurun_out/r2g_pytest.txt
( time python bench.py --impl reference ) > gpurun_out/r2g_bench_reference_n1.json 2> gpurun_out/r2g_bench_reference.err; tail -c 300 gpurun_out/r2g_bench_reference_n1.json; tail -3 gpurun_out/r2g_bench_reference.err
( time python bench.py ) > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2g_bench_n1.json").read().strip().splitlines()[-1])
print("HEADLINE", j["value"], j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "peak", j["roofline"]["peak"], "ref", j.get("reference_gpu", {}).get("speedup_kernel"))
for k, c in j.get("configs", {}).items():
    print(k, "value %.4e" % c["value"], "kernel_ms", c.get("roofline", {}).get("kernel_ms"), "frac", c.get("roofline", {}).get("frac"), "vs ref", c.get("reference_gpu", {}).get("speedup_kernel"))
PY
