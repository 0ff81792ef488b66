#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -k "cpp_multi_gpu or pendulum or fast_fma or opt_in" > gpurun_out/r2r_pytest.log 2>&1
tail -5 gpurun_out/r2r_pytest.log
timeout 600 ilqr_b200/host/_build/bench_batch 8192 200 3 2 | tee gpurun_out/r2r_bench_batch_2gpu.json
timeout 600 ilqr_b200/host/_build/bench_batch 262144 200 2 1 | tee gpurun_out/r2r_bench_batch_2gpu_cfg5.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2r_bench_n2.json').read().strip().split("\n")[-1])
    print("torchrun N=2", d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])
except Exception as e:
    print("N=2 bench failed", e); print(open('gpurun_out/r2r_bench_n2.err').read()[-2000:])
PY
