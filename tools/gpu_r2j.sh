#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/r2j_pytest_all.log
tail -12 gpurun_out/r2j_pytest_all.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r2j_$tool.log 2>&1
  echo "$tool: $(grep -c 'path ' gpurun_out/r2j_$tool.log) configs; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2j_$tool.log | tail -1)"
done
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2j_bench.json'))
    print("bench value %.0f e2e %.0f ms %.2f launches %d frac %.4f" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac']))
    print("secondary", json.dumps(d['roofline']['secondary'])[:600])
    print("cpu", json.dumps(d.get('cpu_baseline'))[:1500])
    for k,v in d.get('extras',{}).items(): print(k, json.dumps(v)[:700])
except Exception as e:
    print("bench FAILED", e); print(open('gpurun_out/r2j_bench.err').read()[-3000:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err
cat gpurun_out/r2j_bench_ref.json | cut -c1-600
