mkdir -p gpurun_out
python -m pytest tests/test_perthread_gpu.py tests/test_fullsize_parity_gpu.py -q -s -k "kellermiksis or c4" > gpurun_out/r2c10_pytest.txt 2>&1; tail -4 gpurun_out/r2c10_pytest.txt; grep "dense\|fullsize\|parity kell" gpurun_out/r2c10_pytest.txt | head -20
MPGOS_B200_LIBRARY=build/variants/libmpgos_b200_bundle1.so python -m pytest tests/test_fullsize_parity_gpu.py -q -s -k "c4" 2>&1 | grep "fullsize\|passed\|failed"
