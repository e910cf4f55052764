python -m pytest tests/test_perthread_gpu.py -x -q -m gpu 2>&1 | tail -3
for lib in "" build/variants/libmpgos_b200_mb2.so build/variants/libmpgos_b200_mb3.so; do
  MPGOS_B200_LIBRARY=$lib python tools/time_perthread.py impact:8 kellermiksis:2 lorenz:1 2>&1 | grep -v Warning
done
