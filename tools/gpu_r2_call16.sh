# round 2, GPU call 16: CUDA-library forms with the three evaluations in one subroutine (stdbundle) against inline calls (default)
mkdir -p gpurun_out
for lib in "" build/variants/libmpgos_b200_stdbundle.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense:2 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2c16_ab.txt
