#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2t_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2t_pytest.log 2>&1
tail -2 gpurun_out/r2t_pytest.log
run cfg2 cfg2 A=1
run cfg5 cfg5 A=1
run cfg5_store cfg5 ILQR_B200_REROLL_MIN=100000000
run cfg5_rr4096 cfg5 ILQR_B200_REROLL_MIN=4096
run cfg3 cfg3 A=1
run cfg3_store cfg3 ILQR_B200_REROLL_MIN=100000000
run cfg4 cfg4 A=1
run cfg4_rr cfg4 ILQR_B200_REROLL_MIN=4096
run cfg2_rr cfg2 ILQR_B200_REROLL_MIN=0
ILQR_B200_REROLL_MIN=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 800 --csv --log-file gpurun_out/r2t_launches_cfg5_rr.csv python tools/profile_solve.py cfg5 > gpurun_out/r2t_ncu.log 2>&1
