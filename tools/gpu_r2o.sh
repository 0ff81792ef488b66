#!/bin/bash
mkdir -p gpurun_out
ILQR_B200_ROWS_GPW=1 ILQR_B200_HANDOVER=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2o_launches_gpw1.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2o_ncu_b.log 2>&1
ILQR_B200_ROWS_GPW=2 ILQR_B200_HANDOVER=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/r2o_launches_gpw2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2o_ncu_b.log 2>&1
