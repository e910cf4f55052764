mkdir -p gpurun_out
TPBS=64,128 REFILLS=0 python tools/time_perthread.py kellermiksis_dense_fm:2 kellermiksis_dense:2 2>&1 | tail -4 | tee gpurun_out/r2c15b_ab.txt
