python -m pytest tests/test_perthread_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   Assert|passed|failed" | cut -c1-200
for lib in "" build/variants/libmpgos_b200_f2f.so build/variants/libmpgos_b200_mb4.so; do
  MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=128 python tools/time_perthread.py impact:8 kellermiksis:2 lorenz:1 lorenz_f32:1 2>&1 | grep -v Warning
done
