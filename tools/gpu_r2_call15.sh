# round 2, GPU call 15: blocks of at least 128 threads for large launches (hint 64 must now time like 128), Keller-Miksis kernels at 96 / 80 registers with the bundled subroutine
mkdir -p gpurun_out
for lib in "" build/variants/libmpgos_b200_km96.so build/variants/libmpgos_b200_km80.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=64,128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense_fm:2 kellermiksis_dense:2 2>&1 | tail -4
done 2>&1 | tee gpurun_out/r2c15_ab.txt
TPBS=64 REFILLS=0 python tools/time_perthread.py lorenz:1 impact:8 duffing:8 lorenz_f32:1 2>&1 | tail -4 | tee -a gpurun_out/r2c15_ab.txt
python -m pytest tests/test_perthread_gpu.py tests/test_fast_math_gpu.py -q -x 2>&1 | tail -2
