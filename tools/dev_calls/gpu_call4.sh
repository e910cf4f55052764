python -m pytest tests/test_perthread_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -3
REFILLS=0,1,4,8,16,32 TPBS=64 python tools/time_perthread.py impact:8 kellermiksis:2 lorenz:1 lorenz_f32:1 2>&1 | grep -v Warning
REFILLS=0,8,32 TPBS=128,256 python tools/time_perthread.py impact:8 kellermiksis:2 lorenz:1 2>&1 | grep -v Warning
