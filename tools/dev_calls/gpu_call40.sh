timeout 300 python -m pytest tests/test_coupled_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^tests|passed|failed|^E   |Error|___" | cut -c1-250 | head -20
DISPATCH=0 TC=1 timeout 120 python tools/time_coupled.py 8192 2>&1 | tail -1
DISPATCH=1 TC=1 timeout 120 python tools/time_coupled.py 8192 2>&1 | tail -1
