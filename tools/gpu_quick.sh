python -m pytest tests/test_perthread_gpu.py tests/test_coupled_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed" | cut -c1-200 | head -5
REFILLS=0 TPBS=128 python tools/time_perthread.py impact:8 kellermiksis:2 lorenz:1 lorenz_f32:1 2>&1 | grep -v Warning
TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
