M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum
mkdir -p gpurun_out
for v in k4 k6; do
  export ILQR_B200_STAGE_MIN=0 ILQR_B200_STAGE_K=${v#k}
  timeout 1200 ncu --metrics $M --clock-control none -c 8000 --csv --log-file gpurun_out/st4_counters_$v.csv python tools/profile_solve.py cfg5 > gpurun_out/st4_c_$v.log 2>&1
  python tools/ncu_solve_summary.py gpurun_out/st4_counters_$v.csv gpurun_out/st4_counters_$v.json > gpurun_out/st4_counters_$v.txt 2>&1
  tail -14 gpurun_out/st4_counters_$v.txt
done
