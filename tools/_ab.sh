mkdir -p gpurun_out
tag=acc1
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_$name.json'))
    print("$name", "%.0f it/s  %.2f ms  fixed15 %.0f  launches %d" % (d['value'], d['ms_per_step'], d['config']['fixed_n_mode']['value'], d['gpu_launches']))
except Exception as e:
    print("$name FAILED", e)
PY
}
for cfg in cfg5 cfg3 cfg4 cfg2; do
  run ${cfg} $cfg
done
run cfg2lock cfg2 ILQR_B200_HANDOVER=0
run cfg2_ord cfg2 ILQR_B200_ORDERED_MIN=0
run cfg4_ord cfg4 ILQR_B200_ORDERED_MIN=0
