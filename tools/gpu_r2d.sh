#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2d_pytest_phase.log 2>&1
tail -3 gpurun_out/r2d_pytest_phase.log
for wm in 6144 2500 1000 0; do
ILQR_B200_WARP_PRE_MAX=$wm timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d_bench_cfg2_$wm.json 2> gpurun_out/r2d_bench_cfg2_$wm.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2d_bench_cfg2_$wm.json'))
print("cfg2 warp_pre_max=$wm", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
done
ILQR_B200_WARP_PRE_MAX=16384 timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2d_bench_cfg4.json 2> gpurun_out/r2d_bench_cfg4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_cfg4.json'))
print("cfg4 (pre_max 16384)", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2d_bench_cfg4b.json 2> gpurun_out/r2d_bench_cfg4b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_cfg4b.json'))
print("cfg4 (default)", d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2d_ncu_b.log 2>&1
