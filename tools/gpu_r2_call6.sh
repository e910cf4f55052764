# round 2, GPU call 6: the bundled fast-math subroutine (accuracy, A/B against three separate calls on C4 and C5), the tutorial programs
# T1/T3/T4/T5/T7 against the checker, one full ncu capture of the C5 fast-math kernel
mkdir -p gpurun_out
( time python -m pytest tests/test_fast_math_gpu.py tests/test_tutorials_run_gpu.py -x -q -s ) > gpurun_out/r2c6_pytest.txt 2>&1; tail -5 gpurun_out/r2c6_pytest.txt; grep -E "^\[|\[tutorial|\[fast_math\]" gpurun_out/r2c6_pytest.txt | head -40
for lib in "" build/variants/libmpgos_b200_nobundle.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense_fm:2 kellermiksis_dense:2 2>&1 | tail -2
  MPGOS_B200_LIBRARY=$lib FM=1 python tools/time_coupled.py 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2c6_ab.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:perblock_kernel -c 2 -o gpurun_out/r2c6_coupled_fm -f env FM=1 python tools/time_coupled.py > gpurun_out/r2c6_ncu.log 2>&1; tail -2 gpurun_out/r2c6_ncu.log
