show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['e2e'])"; }
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu | show chunks4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --e2e-chunks 8 | show chunks8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --e2e-chunks 2 | show chunks2
python bench.py --workload coupled --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu | show coupled
