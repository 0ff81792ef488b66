mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "rc=$?"
tail -5 gpurun_out/n2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/n2_bench.json"))
print(d["value"], d["ms_per_step"], d["per_rank"])
PY
