# several units per thread in the block kernel: the new cases, the coupled suite, C5 timing (must be unchanged: 13.6 / 17.1 ms)
mkdir -p gpurun_out
( time python -m pytest tests/test_coupled_gpu.py tests/test_fullsize_parity_gpu.py tests/test_fast_math_gpu.py tests/test_sharded_solver.py tests/test_dropin_gpu.py -m gpu -q -s ) > gpurun_out/upt_pytest.txt 2>&1; tail -5 gpurun_out/upt_pytest.txt; grep "ring .* units\|C5" gpurun_out/upt_pytest.txt
FM=1 python tools/time_coupled.py 2>&1 | tail -1
python tools/time_coupled.py 2>&1 | tail -1
F32=1 python tools/time_coupled.py 2>&1 | tail -1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_coupled_gpu.py -m gpu -q -k "more_than_1024" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | head
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_coupled_gpu.py -m gpu -q -k "more_than_1024" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed" | head
