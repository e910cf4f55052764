python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-400 | head -20
for p in f64 f32; do python bench.py --workload coupled --precision $p --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('coupled $p value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), 'kernel_ms %.3f'%d['roofline']['kernel_ms'], d['e2e'])"; done
