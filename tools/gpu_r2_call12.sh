# round 2, GPU call 12: the whole GPU suite without -x (every gate that was tightened in round 2 reports its margin)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/r2c12_pytest.txt 2>&1; tail -15 gpurun_out/r2c12_pytest.txt; grep -E "^\[|^\.\[|^F\[|\]\s" gpurun_out/r2c12_pytest.txt | grep "parity\|fullsize\|dense\|accessor" | head -60
