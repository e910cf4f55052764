# development: compare A/B builds of the library (build/variants/lib*.so) on the headline workload
for lib in "" $(ls build/variants/lib*.so 2>/dev/null); do
  for refill in 0 1; do
    for tpb in 64 128; do
      MPGOS_B200_LIBRARY=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu --lane-refill $refill --threads-per-block $tpb "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('lib=%-28s refill=$refill tpb=$tpb value %.3e kernel_ms %.3f frac %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
    done
  done
done
