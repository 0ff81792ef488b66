#!/bin/bash
# A/B of two builds of the library: tools/gpu_ab.sh <tag> <libA> <libB> -- runs the bench configs with each
mkdir -p gpurun_out
tag=$1; shift
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
for lib in "$@"; do
  L=$PWD/ilqr_b200/$lib
  run cfg2_$lib cfg2 ILQR_B200_LIB=$L
  run cfg2lock_$lib cfg2 ILQR_B200_LIB=$L ILQR_B200_HANDOVER=0
  run cfg5_$lib cfg5 ILQR_B200_LIB=$L
  run cfg3_$lib cfg3 ILQR_B200_LIB=$L
  run cfg4_$lib cfg4 ILQR_B200_LIB=$L
done
