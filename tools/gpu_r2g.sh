#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2g_pytest_phase.log 2>&1
tail -3 gpurun_out/r2g_pytest_phase.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2g_bench_cfg2.json 2> gpurun_out/r2g_bench_cfg2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_cfg2.json'))
print("cfg2", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2g_bench_cfg4.json 2> gpurun_out/r2g_bench_cfg4.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_cfg4.json'))
print("cfg4", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
ILQR_B200_ROWS_MAX=1000000 timeout 600 python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2g_bench_cfg5.json 2> gpurun_out/r2g_bench_cfg5.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_cfg5.json'))
print("cfg5 rows only", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2g_ncu_b.log 2>&1
