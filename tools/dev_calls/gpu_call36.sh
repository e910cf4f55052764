python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-400 | head -20
for d in 0 1; do for r in "" 1; do
DISPATCH=$d REVERSE=$r REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz:2 impact:8 kellermiksis:2 2>&1 | grep "lib="
done; done
