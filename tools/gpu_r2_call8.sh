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
