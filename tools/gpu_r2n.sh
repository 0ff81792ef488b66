#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2n_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
run cfg2 cfg2 A=1
run cfg2_lock cfg2 ILQR_B200_HANDOVER=0
run cfg5 cfg5 A=1
run cfg3 cfg3 A=1
ILQR_B200_HANDOVER=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2n_ncu_b.log 2>&1
ILQR_B200_HANDOVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:backward_rows -s 305 -c 1 -o gpurun_out/r2n_rows_bulk python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2n_ncu.log 2>&1
ILQR_B200_HANDOVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_rollout -s 305 -c 1 -o gpurun_out/r2n_roll_bulk python bench.py --steps 1 --warmup 3 --no-cpu --no-extras >> gpurun_out/r2n_ncu.log 2>&1
ILQR_B200_HANDOVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_sweep -s 305 -c 1 -o gpurun_out/r2n_sweep_bulk python bench.py --steps 1 --warmup 3 --no-cpu --no-extras >> gpurun_out/r2n_ncu.log 2>&1
