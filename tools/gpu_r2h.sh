#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2h_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2h_pytest_phase.log 2>&1
tail -2 gpurun_out/r2h_pytest_phase.log
run gpw4 ILQR_B200_ROWS_GPW=4
run gpw2 ILQR_B200_ROWS_GPW=2
run gpw1 ILQR_B200_ROWS_GPW=1
run gpw4_ho2400 ILQR_B200_ROWS_GPW=4 ILQR_B200_HANDOVER=2400 ILQR_B200_CHECK_EVERY=4
run gpw4_ho3200 ILQR_B200_ROWS_GPW=4 ILQR_B200_HANDOVER=3200 ILQR_B200_CHECK_EVERY=4
run gpw4_ho1200 ILQR_B200_ROWS_GPW=4 ILQR_B200_HANDOVER=1200 ILQR_B200_CHECK_EVERY=4
run gpw2_ho2400 ILQR_B200_ROWS_GPW=2 ILQR_B200_HANDOVER=2400 ILQR_B200_CHECK_EVERY=4
run gpw1_ho2400 ILQR_B200_ROWS_GPW=1 ILQR_B200_HANDOVER=2400 ILQR_B200_CHECK_EVERY=4
run ho4096 ILQR_B200_HANDOVER=4096
