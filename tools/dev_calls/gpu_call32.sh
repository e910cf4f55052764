for w in coupled impact kellermiksis_dense duffing; do
  python bench.py --workload $w --precision f32 --steps 3 --warmup 3 $( [ $w = impact ] && echo --phases 8 ) $( [ $w = kellermiksis_dense ] && echo --phases 2 ) $( [ $w = duffing ] && echo --phases 8 ) > gpurun_out/bench_f32_$w.json 2> gpurun_out/bench_f32_$w.err
  tail -1 gpurun_out/bench_f32_$w.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('$w', d['metric'], 'value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), 'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind'))
except Exception as e: print('$w FAILED', e); print(open('gpurun_out/bench_f32_$w.err').read()[-1500:])"
done
