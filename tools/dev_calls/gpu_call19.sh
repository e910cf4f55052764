python -m pytest tests/test_coupled_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed" | cut -c1-200 | head
for tc in 0 1 2; do TC=$tc python tools/time_coupled.py 8192 2>&1 | tail -1; done
TC=2 python tools/time_coupled.py 65536 2>&1 | tail -1
