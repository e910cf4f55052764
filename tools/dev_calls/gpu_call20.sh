# round-1 re-entry: full GPU test-suite on the new RKCK45 arithmetic (increment-form stages, FP32-seeded single FP64 Newton step),
# then A/B timing of the variants
python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-300 | head -20
for lib in "" build/variants/libmpgos_b200_old.so build/variants/libmpgos_b200_stage.so build/variants/libmpgos_b200_newton.so build/variants/libmpgos_b200_r128.so; do
  MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis:2 lorenz_f32:1 2>&1 | grep "lib="
done
