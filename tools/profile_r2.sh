#!/bin/bash
# round-2 profile pass (run under gpurun): launch lists with DRAM / instruction counters for one solve of each
# configuration, and ncu --set full captures of the top kernels.  Outputs under gpurun_out/r2z_*.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum
# the command of the profiling guide: every launch of a short bench run
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2z_b.log 2>&1
for cfg in cfg2 cfg4 cfg3 cfg5; do
  timeout 1500 ncu --metrics $M --clock-control none -c 6000 --csv --log-file gpurun_out/r2z_counters_$cfg.csv python tools/profile_solve.py $cfg > gpurun_out/r2z_c_$cfg.log 2>&1
  python tools/ncu_solve_summary.py gpurun_out/r2z_counters_$cfg.csv gpurun_out/r2z_counters_$cfg.json > gpurun_out/r2z_counters_$cfg.txt 2>&1
  tail -12 gpurun_out/r2z_counters_$cfg.txt
done
for k in backward_rows phase_rollout phase_sweep phase_accept ilqr_warp_kernel; do
  skip=20; [ $k = ilqr_warp_kernel ] && skip=4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2z_$k python tools/profile_solve.py cfg2 > gpurun_out/r2z_full_$k.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_backward_kernel -s 20 -c 1 -o gpurun_out/r2z_phase_backward_thread python tools/profile_solve.py cfg5 > gpurun_out/r2z_full_thread.log 2>&1
ls -la gpurun_out/r2z_*.ncu-rep
