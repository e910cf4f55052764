python -m pytest tests/test_perthread_gpu.py tests/test_dropin_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   Assert|passed|failed" | cut -c1-200
REFILLS=0 TPBS=128 python tools/time_perthread.py impact:8 kellermiksis:2 lorenz:1 lorenz_f32:1 2>&1 | grep -v Warning
