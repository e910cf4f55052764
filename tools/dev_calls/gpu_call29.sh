python -m pytest tests -x -q -m gpu --tb=short 2>&1 | grep -E "^E   |passed|failed|Error" | cut -c1-300 | head -20
REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis:2 kellermiksis_dense:2 lorenz_f32:1 duffing:4 2>&1 | grep "lib="
ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum --clock-control none -k regex:perthread_kernel -c 1 python tools/time_perthread.py lorenz:1 2>&1 | grep -E "inst_executed|duration"
