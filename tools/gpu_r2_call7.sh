# round 2, GPU call 7: fast-math bundle with two reciprocals (accuracy, A/B on C4 and C5 against the three-function bundle and separate calls),
# C5 at 768 / 1024 resident threads per SM (80 / 64 registers), tutorial programs
mkdir -p gpurun_out
( time python -m pytest tests/test_fast_math_gpu.py tests/test_tutorials_run_gpu.py -x -q -s ) > gpurun_out/r2c7_pytest.txt 2>&1; tail -5 gpurun_out/r2c7_pytest.txt; grep -E "\[tutorial|\[fast_math\] (b|K)" gpurun_out/r2c7_pytest.txt | head -40
for lib in "" build/variants/libmpgos_b200_bundle1.so build/variants/libmpgos_b200_nobundle.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense_fm:2 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib FM=1 python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2c7_ab.txt
for lib in build/variants/libmpgos_b200_s4t768.so build/variants/libmpgos_b200_s4t1024.so; do
  MPGOS_B200_LIBRARY=$lib FM=1 python tools/time_coupled.py 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee -a gpurun_out/r2c7_ab.txt
