# round 2, GPU call 8: fast-math forms with the numerator summed as a tree and two reciprocals side by side, A/B against the plain bundle
mkdir -p gpurun_out
( time python -m pytest tests/test_fast_math_gpu.py -x -q -s ) > gpurun_out/r2c8_pytest.txt 2>&1; tail -4 gpurun_out/r2c8_pytest.txt; grep -E "\[fast_math\] (b|K|r)" gpurun_out/r2c8_pytest.txt | head -40
for lib in "" build/variants/libmpgos_b200_bundle1.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense_fm:2 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib FM=1 python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2c8_ab.txt
# four DMMA accumulator chains; two blocks of four systems per SM with four chains
for lib in build/variants/libmpgos_b200_dmma4.so build/variants/libmpgos_b200_s4dmma4.so; do
  MPGOS_B200_LIBRARY=$lib FM=1 python tools/time_coupled.py 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee -a gpurun_out/r2c8_ab.txt
python - <<'PY'
import importlib, os, sys
import numpy as np
sys.path[:0] = [os.getcwd()]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
rng = np.random.default_rng(5)
for scale in (1.0, 10.0, 1e3, 1e5, 1e8, 1e15, 1e300):
    x = rng.uniform(-scale, scale, 2_000_000)
    s, c, s2, c2 = pkg.fast_math_probe(7, x)
    print(f"CUDA sin/cos vs sincos, |x| <= {scale:g}: different sin {int((s != s2).sum())}, different cos {int((c != c2).sum())} of {x.size}")
x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, np.pi, np.pi / 2, 105414357.0, 1e22])
s, c, s2, c2 = pkg.fast_math_probe(7, x)
print("special", np.array_equal(s, s2, equal_nan=True), np.array_equal(c, c2, equal_nan=True), np.array_equal(np.signbit(s), np.signbit(s2)))
PY
# CUDA-library forms with one ::sincos instead of sin + cos
for lib in "" build/variants/libmpgos_b200_stdsincos.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense:2 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee -a gpurun_out/r2c8_ab.txt
