python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^tests|passed|failed|^E   |Error|___" | cut -c1-250 | head -20
for d in 0 1; do for r in "" 1; do
PHASES=2 DISPATCH=$d REVERSE=$r TC=1 python tools/time_coupled.py 8192 2>&1 | tail -1
done; done
