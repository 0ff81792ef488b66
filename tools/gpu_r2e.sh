#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2e_pytest_phase.log 2>&1
tail -5 gpurun_out/r2e_pytest_phase.log
for rm in 24576 0; do
ILQR_B200_ROWS_MAX=$rm timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2e_bench_cfg2_$rm.json 2> gpurun_out/r2e_bench_cfg2_$rm.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_cfg2_$rm.json'))
print("cfg2 rows_max=$rm", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
done
for rm in 24576 0; do
ILQR_B200_ROWS_MAX=$rm timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2e_bench_cfg4_$rm.json 2> gpurun_out/r2e_bench_cfg4_$rm.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_cfg4_$rm.json'))
print("cfg4 rows_max=$rm", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
done
for rm in 1000000 24576; do
ILQR_B200_ROWS_MAX=$rm timeout 600 python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2e_bench_cfg5_$rm.json 2> gpurun_out/r2e_bench_cfg5_$rm.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_cfg5_$rm.json'))
print("cfg5 rows_max=$rm", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2e_ncu_b.log 2>&1
