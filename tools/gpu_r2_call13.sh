# round 2, GPU call 13 (2 GPUs): the tutorial programs once more (T8 on devices 0 and 1, T5 and the bubble tutorial with the fast-math redirect)
mkdir -p gpurun_out
( time python -m pytest tests/test_tutorials_run_gpu.py tests/test_sharded_solver.py -m gpu -q -s ) > gpurun_out/r2c13_pytest.txt 2>&1; tail -6 gpurun_out/r2c13_pytest.txt; grep "tutorial" gpurun_out/r2c13_pytest.txt | head -30
