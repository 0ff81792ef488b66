#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/tmp_$name.json 2> gpurun_out/tmp_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/tmp_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
run base cfg2 A=1
run gpw2 cfg2 ILQR_B200_ROWS_GPW=2
run gpw2_lock cfg2 ILQR_B200_ROWS_GPW=2 ILQR_B200_HANDOVER=0
run gpw4_lock cfg2 ILQR_B200_HANDOVER=0
run ho2800 cfg2 ILQR_B200_HANDOVER=2800
run ho3600 cfg2 ILQR_B200_HANDOVER=3600
run ho3600c2 cfg2 ILQR_B200_HANDOVER=3600 ILQR_B200_CHECK_EVERY=2
run ho3200c2 cfg2 ILQR_B200_HANDOVER=3200 ILQR_B200_CHECK_EVERY=2
run ho3900c1 cfg2 ILQR_B200_HANDOVER=3900 ILQR_B200_CHECK_EVERY=1
run cfg4_gpw2 cfg4 ILQR_B200_ROWS_GPW=2
run cfg4_base cfg4 A=1
