python -m pytest tests/test_coupled_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-400 | head -20
F32=1 TC=0 python tools/time_coupled.py 8192 2>&1 | tail -1
F32=1 TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
UPS=32 F32=1 TC=0 python tools/time_coupled.py 8192 2>&1 | tail -1
UPS=32 F32=1 TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
