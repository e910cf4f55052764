for lib in "" build/variants/libmpgos_b200_r64.so build/variants/libmpgos_b200_r72.so; do
MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64,128 python tools/time_perthread.py lorenz:1 impact:8 2>&1 | grep "lib="
done
