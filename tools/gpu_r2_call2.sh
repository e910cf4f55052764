# round 2, GPU call 2: full GPU suite, sharded solver, variants, dense-output store efficiency, ncu summaries
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2c2_pytest.txt 2>&1; tail -6 gpurun_out/r2c2_pytest.txt; grep "fullsize" gpurun_out/r2c2_pytest.txt
for lib in "" ; do
  MPGOS_B200_LIBRARY=$lib REFILLS=0 TPBS=64,128 python tools/time_perthread.py lorenz:1 impact:8 kellermiksis_dense:2 duffing:8 lorenz_f32:1 2>&1 | grep -v "^$" | cut -c1-170
done
M="l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_write.sum,dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum"
for c in kellermiksis_dense duffing_events impact_dense; do
  ncu --metrics $M --clock-control none -k regex:perthread_kernel --csv --log-file gpurun_out/r2c2_dense_$c.csv python tools/probe_dense_stores.py $c > gpurun_out/r2c2_dense_$c.txt 2>&1; tail -1 gpurun_out/r2c2_dense_$c.txt
done
ncu --metrics $M --clock-control none -k regex:perblock_kernel --csv --log-file gpurun_out/r2c2_dense_coupled_ring.csv python tools/probe_dense_stores.py coupled_ring > gpurun_out/r2c2_dense_coupled_ring.txt 2>&1; tail -1 gpurun_out/r2c2_dense_coupled_ring.txt
for w in lorenz impact kellermiksis_dense; do
  ncu --set full --clock-control none --import-source on -k regex:perthread_kernel -s 3 -c 1 -f -o gpurun_out/r2c2_prof_$w python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --configs none > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:perblock_kernel -s 3 -c 1 -f -o gpurun_out/r2c2_prof_coupled python bench.py --workload coupled --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --configs none > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
