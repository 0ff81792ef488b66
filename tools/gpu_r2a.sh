#!/bin/bash
# round-2 first GPU pass: engine equivalence, bench of the phase engine, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_engine" > gpurun_out/r2a_pytest_phase.log 2>&1
echo "phase tests rc=$?" >> gpurun_out/r2a_pytest_phase.log
tail -5 gpurun_out/r2a_pytest_phase.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
cat gpurun_out/r2a_bench_cfg2.json
ILQR_B200_ENGINE=warp timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg2_warp.json 2> gpurun_out/r2a_bench_cfg2_warp.err
cat gpurun_out/r2a_bench_cfg2_warp.json
timeout 600 python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg5.json 2> gpurun_out/r2a_bench_cfg5.err
cat gpurun_out/r2a_bench_cfg5.json
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err
cat gpurun_out/r2a_bench_cfg4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2a_ncu_b.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/r2a_pytest_all.log
tail -15 gpurun_out/r2a_pytest_all.log
