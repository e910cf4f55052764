python -m pytest tests/test_statistics_gpu.py -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-300 | head -20
REFILLS=0 TPBS=32,64,128 python tools/time_perthread.py duffing:4 2>&1 | grep "lib="
