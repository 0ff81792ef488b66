#!/bin/bash
# the driver's round-end sequence: GPU tests, smoke, sanitizers, bench (ours + reference arm)
mkdir -p gpurun_out
tag=${1:-final}
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/${tag}_pytest_all.log
tail -6 gpurun_out/${tag}_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/${tag}_$tool.log 2>&1
  echo "$tool: $(grep -c 'path ' gpurun_out/${tag}_$tool.log) configs; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/${tag}_$tool.log | tail -1)"
done
timeout 900 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
cut -c1-300 gpurun_out/${tag}_bench_ref.json
timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_bench.json'))
    print("bench value %.0f e2e %.0f ms %.2f launches %d frac %.4f traffic %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic']))
    print("secondary", {k:v for k,v in d['roofline']['secondary'].items() if k not in ('measured_fp64_ops_per_s','note','bound')})
    print("cpu", json.dumps(d.get('cpu_baseline'))[:1200])
    for k,v in d.get('extras',{}).items(): print(k, json.dumps(v)[:600])
    print("clocks", d['clocks'])
except Exception as e:
    print("bench FAILED", e); print(open('gpurun_out/${tag}_bench.err').read()[-3000:])
PY
