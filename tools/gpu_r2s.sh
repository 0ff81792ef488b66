#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2s_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine or bit_exact" > gpurun_out/r2s_pytest.log 2>&1
tail -2 gpurun_out/r2s_pytest.log
run cfg2 cfg2 A=1
run cfg2_lock cfg2 ILQR_B200_HANDOVER=0
run cfg2_ho2400 cfg2 ILQR_B200_HANDOVER=2400
run cfg2_ho1600 cfg2 ILQR_B200_HANDOVER=1600
run cfg2_warp cfg2 ILQR_B200_ENGINE=warp
run cfg2_thread cfg2 ILQR_B200_ROWS_MAX=0 ILQR_B200_HANDOVER=0
run cfg4 cfg4 A=1
run cfg5 cfg5 A=1
run cfg5_rows cfg5 ILQR_B200_ROWS_MAX=1000000
run cfg3 cfg3 A=1
ILQR_B200_HANDOVER=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2s_ncu_b.log 2>&1
