#!/bin/bash
# BASELINE configs[4] on the 8 GPUs of one box: bash tools/gpu_n8.sh   (run under gpurun --gpus 8)
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config cfg5 --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/n${N}_cfg5.json 2> gpurun_out/n${N}_cfg5.err
echo "rc=$?"
python - <<PY
import json
t=open("gpurun_out/n${N}_cfg5.json").read()
d=json.loads([l for l in t.split("\n") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("per_rank"))
PY
