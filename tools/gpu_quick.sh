python -m pytest tests/test_coupled_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed" | cut -c1-300 | head -8
UPS=128 TC=0 python tools/time_coupled.py 4096 2>&1 | tail -1
UPS=128 TC=1 REFTAG=bubble_ensemble_u128 python tools/time_coupled.py 4096 2>&1 | grep -vE "^ |^$|memory|Object|THREAD|GLOBAL|SHARED|\-\-\-" | tail -4
