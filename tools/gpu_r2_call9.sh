# round 2, GPU call 9: CUDA-library forms of the built-in bubble systems with one ::sincos instead of sin + cos (bit-identical, see call 8)
mkdir -p gpurun_out
for lib in "" build/variants/libmpgos_b200_stdsincos.so; do
  MPGOS_B200_LIBRARY=$lib TPBS=128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense:2 2>&1 | tail -1
  MPGOS_B200_LIBRARY=$lib python tools/time_coupled.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2c9_ab.txt
