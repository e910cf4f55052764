for lib in "" build/variants/libmpgos_b200_rev.so build/variants/libmpgos_b200_mix.so; do
MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis:2 lorenz_f32:1 2>&1 | grep "lib="
done
