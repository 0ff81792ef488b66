#!/bin/bash
# staged line search (ILQR_B200_STAGE_MIN / _K, ilqr_phase_launch.cuh): off against a few settings, on the large configs
mkdir -p gpurun_out
tag=${1:-stage}
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 900 python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
OFF=4611686018427387904
for cfg in cfg5 cfg3; do
  run ${cfg}_off $cfg ILQR_B200_STAGE_MIN=$OFF
  for k in 2 4 6; do
    run ${cfg}_k${k} $cfg ILQR_B200_STAGE_MIN=0 ILQR_B200_STAGE_K=$k
  done
done
run cfg5_k4_reroll cfg5 ILQR_B200_STAGE_MIN=0 ILQR_B200_STAGE_K=4 ILQR_B200_REROLL_MIN=0
run cfg5_reroll cfg5 ILQR_B200_STAGE_MIN=$OFF ILQR_B200_REROLL_MIN=0
run cfg5_default cfg5
for cfg in cfg4 cfg2; do
  run ${cfg}_off $cfg ILQR_B200_STAGE_MIN=$OFF
  run ${cfg}_k4 $cfg ILQR_B200_STAGE_MIN=0 ILQR_B200_STAGE_K=4
done
