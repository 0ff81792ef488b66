#!/bin/bash
mkdir -p gpurun_out
run() { # name, config, env...
  name=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2i_pytest_phase.log 2>&1
tail -2 gpurun_out/r2i_pytest_phase.log
run cfg2_lock cfg2 A=1
run cfg2_ho3200 cfg2 ILQR_B200_HANDOVER=3200 ILQR_B200_CHECK_EVERY=4
run cfg2_ho2800 cfg2 ILQR_B200_HANDOVER=2800 ILQR_B200_CHECK_EVERY=4
run cfg2_ho3600 cfg2 ILQR_B200_HANDOVER=3600 ILQR_B200_CHECK_EVERY=2
run cfg4_lock cfg4 A=1
run cfg4_ho3200 cfg4 ILQR_B200_HANDOVER=3200 ILQR_B200_CHECK_EVERY=4
run cfg4_ho6000 cfg4 ILQR_B200_HANDOVER=6000 ILQR_B200_CHECK_EVERY=4
run cfg5_lock cfg5 A=1
run cfg5_ho3200 cfg5 ILQR_B200_HANDOVER=3200
run cfg3_lock cfg3 A=1
run cfg3_ho3200 cfg3 ILQR_B200_HANDOVER=3200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2i_ncu_b.log 2>&1
