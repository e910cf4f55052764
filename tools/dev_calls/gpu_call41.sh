cat > /tmp/san_only.sh <<'SH'
sed -n '/^cat > \/tmp\/san_case.py/,/^PY$/p' tools/gpu_sanitize.sh | sed '1d;$d' > /tmp/san_case.py
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py dispatch > gpurun_out/sanitize_${tool}_dispatch.txt 2>&1
  echo "== $tool dispatch: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_${tool}_dispatch.txt | sort | uniq -c | head -5; grep -cE "^(bubble|ring|impact)" gpurun_out/sanitize_${tool}_dispatch.txt
done
SH
bash /tmp/san_only.sh
python bench.py --workload coupled --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_coupled.json
python -c "
import json
d=json.loads(open('gpurun_out/final_coupled.json').read()); print('%.4e'%d['value'], d['roofline']['kernel_ms'], d['reference_gpu']['speedup_kernel'], d['gpu_launches'])"
