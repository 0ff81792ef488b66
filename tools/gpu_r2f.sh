#!/bin/bash
mkdir -p gpurun_out
# one tail launch (trip ~80 of the 4th solve) and one bulk launch of the rows kernel, with source counters
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backward_rows -s 385 -c 1 -o gpurun_out/r2f_rows_tail python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2f_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backward_rows -s 305 -c 1 -o gpurun_out/r2f_rows_bulk python bench.py --steps 1 --warmup 3 --no-cpu >> gpurun_out/r2f_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_rollout -s 385 -c 1 -o gpurun_out/r2f_roll_tail python bench.py --steps 1 --warmup 3 --no-cpu >> gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
