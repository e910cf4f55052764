python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^tests|passed|failed|^E   |Error|___" | cut -c1-250 | head -20
for lib in "" build/variants/libmpgos_b200_nopack.so; do
MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz_f32:1 lorenz:1 2>&1 | grep "lib="
for w in impact kellermiksis_dense; do
  MPGOS_B200_LIBRARY=$lib python bench.py --workload $w --precision f32 --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu $( [ $w = impact ] && echo --phases 8 ) $( [ $w = kellermiksis_dense ] && echo --phases 2 ) 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('lib=$lib $w f32 value %.3e kernel_ms %.4f'%(d['value'], d['roofline']['kernel_ms']))"
done; done
