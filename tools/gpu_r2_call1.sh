# round 2, GPU call 1: GPU test suite (incl. the new full-size parity tests), the bench line with its configs object, A/B timing of library variants
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2c1_pytest.txt 2>&1; tail -5 gpurun_out/r2c1_pytest.txt; grep "fullsize" gpurun_out/r2c1_pytest.txt
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; tail -4 gpurun_out/r2c1_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c1_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("HEADLINE value %.4e kernel_ms %.4f frac %.4f peak %.2f probes %s e2e %.4e refgpu %s" % (d["value"], r["kernel_ms"], r["frac"], r["peak"], r["peak_probes"], d["e2e"]["value"], {k:d.get("reference_gpu",{}).get(k) for k in ("kernel_ms","speedup_kernel","best_setting")}))
    print("  refgpu settings", d.get("reference_gpu",{}).get("kernel_ms_by_setting"))
    for k,c in d.get("configs",{}).items():
        if "error" in c: print(k, c); continue
        r=c["roofline"]; g=c.get("reference_gpu",{})
        print(k, "value %.4e kernel_ms %.4f frac %.4f e2e %.4e refgpu x%s (%s) cpu %s" % (c["value"], r["kernel_ms"], r["frac"], c["e2e"]["value"], g.get("speedup_kernel"), g.get("best_setting"), c.get("cpu_baseline",{}).get("value")))
        print("   ", g.get("kernel_ms_by_setting"))
except Exception as e: print("bench unreadable", e)
PY
for lib in "" build/variants/libmpgos_b200_r1.so build/variants/libmpgos_b200_accept0.so build/variants/libmpgos_b200_regs96.so build/variants/libmpgos_b200_regs128.so; do
  MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64,128 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis_dense:2 duffing:8 lorenz_f32:1 2>&1 | grep -v "^$" | cut -c1-170
done
