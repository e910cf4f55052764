# usage: bash tools/build_variant.sh TAG "extra nvcc flags"   -> build/variants/libmpgos_b200_TAG.so  (A/B builds for tools/sweep_variants.sh)
set -e
TAG=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/massively-parallel-gpu-ode-solver_b200/csrc
OUT=$ROOT/build/variants/$TAG
mkdir -p $OUT
for s in runtime runtime_coupled builtin_lorenz builtin_duffing builtin_impact builtin_kellermiksis builtin_coupled; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v $EXTRA -c $SRC/$s.cu -o $OUT/$s.o 2> $OUT/$s.ptxas.log &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/build/variants/libmpgos_b200_$TAG.so $OUT/*.o
grep -h -A2 "perthread_kernel.*\(Impact\|KellerMiksis\)" $OUT/builtin_impact.ptxas.log $OUT/builtin_kellermiksis.ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    : //' | paste - - | cut -c1-160
