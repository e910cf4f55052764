TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
MPGOS_B200_LIBRARY=build/variants/libmpgos_b200_mma4.so TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
MPGOS_B200_LIBRARY=build/variants/libmpgos_b200_mma4.so python -m pytest tests/test_coupled_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-300 | head
UPS=128 TC=1 python tools/time_coupled.py 4096 2>&1 | tail -1
UPS=32 TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
MPGOS_B200_LIBRARY=build/variants/libmpgos_b200_mma4.so UPS=32 TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
