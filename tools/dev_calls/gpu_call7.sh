python -m pytest tests/test_dropin_gpu.py -x -q -m gpu -k coupled 2>&1 | grep -E "^E|assert|Error|passed|failed" | cut -c1-300 | head -30
export LD_LIBRARY_PATH=$PWD/massively-parallel-gpu-ode-solver_b200:$LD_LIBRARY_PATH
cd build/tut && mkdir -p ring bubble
(cd ring && timeout 120 ../RingOfDuffingOscillators.exe > log.txt 2>&1; echo ring rc=$?; tail -3 log.txt; ls | head; head -c 600 RingOfDuffingOscillators.txt 2>/dev/null)
(cd bubble && timeout 300 ../BubbleEnsemble.exe > log.txt 2>&1; echo bubble rc=$?; tail -4 log.txt; ls | head)
