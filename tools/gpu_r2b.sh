#!/bin/bash
# profile ONE backward-phase launch (and one rollout / sweep launch) of the phase engine with source-level counters
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_backward -s 12 -c 1 -o gpurun_out/r2b_backward python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_rollout -s 12 -c 1 -o gpurun_out/r2b_rollout python bench.py --steps 1 --warmup 3 --no-cpu >> gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
ls -la gpurun_out/*.ncu-rep
