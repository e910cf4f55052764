REFILLS=0,1,8,32 TPBS=32,64 python tools/time_perthread.py impact:32 2>&1 | grep -v Warning
REFILLS=0 TPBS=32 python tools/time_perthread.py kellermiksis:2 lorenz:1 2>&1 | grep -v Warning
