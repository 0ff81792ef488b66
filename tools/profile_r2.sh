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
  rm -f gpurun_out/r2z_counters_$cfg.csv
done
# ncu --set full captures, summarised on the box (the reports themselves are ~8 MB each: over gpurun's 64 MiB return limit)
summarise() { rep=$1; out=$2
  ncu -i $rep --page raw | sed -e 's/^ *//' | grep -E '^[a-z0-9_]+__' > gpurun_out/${out}_raw.txt 2>/dev/null
  ncu -i $rep --page source --print-source cuda,sass --csv > /tmp/_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/_src.csv 40 > gpurun_out/${out}_lines.txt 2>&1
  rm -f $rep /tmp/_src.csv
}
for k in backward_rows phase_rollout phase_sweep phase_accept phase_compact ilqr_warp_kernel; do
  skip=20; [ $k = ilqr_warp_kernel ] && skip=3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/r2z_$k python tools/profile_solve.py cfg2 > gpurun_out/r2z_full_$k.log 2>&1
  summarise /tmp/r2z_$k.ncu-rep r2z_$k
done
# the large-batch regime (configs[4] shard, third trip: near-full active set)
for k in phase_backward_kernel phase_rollout phase_accept phase_sweep phase_compact; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/r2z_cfg5_$k python tools/profile_solve.py cfg5 > gpurun_out/r2z_full_cfg5_$k.log 2>&1
  summarise /tmp/r2z_cfg5_$k.ncu-rep r2z_cfg5_$k
done
ls -la gpurun_out/ | head -60
du -sh gpurun_out
