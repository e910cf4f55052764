python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-300 | head -20
for lib in "" build/variants/libmpgos_b200_norow.so build/variants/libmpgos_b200_r96.so build/variants/libmpgos_b200_norow96.so build/variants/libmpgos_b200_r72.so; do
  MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis:2 kellermiksis_dense:2 lorenz_f32:1 duffing:4 2>&1 | grep "lib="
done
