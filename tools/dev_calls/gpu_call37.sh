for d in 1 2; do for r in "" 1; do
DISPATCH=$d REVERSE=$r REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis:2 kellermiksis:1 2>&1 | grep "lib="
done; done
