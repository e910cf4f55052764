python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed" | cut -c1-200 | head
python bench.py --steps 10 --e2e-chunks 4 --no-cpu-baseline --no-reference-gpu | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('chunks4 value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
python bench.py --steps 10 --e2e-chunks 8 --no-cpu-baseline --no-reference-gpu | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('chunks8 value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
python bench.py --steps 10 --e2e-chunks 1 --no-cpu-baseline --no-reference-gpu | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('chunks1 value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
python bench.py --workload kellermiksis_dense --phases 2 --steps 3 --no-cpu-baseline --no-reference-gpu | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['dense_output'], d['e2e'])"
