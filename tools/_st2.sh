M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum
mkdir -p gpurun_out
timeout 1200 ncu --metrics $M --clock-control none -c 8000 --csv --log-file gpurun_out/ord2_counters_cfg5.csv python tools/profile_solve.py cfg5 > gpurun_out/ord2_c.log 2>&1
python tools/ncu_solve_summary.py gpurun_out/ord2_counters_cfg5.csv gpurun_out/ord2_counters_cfg5.json > gpurun_out/ord2_counters_cfg5.txt 2>&1
tail -14 gpurun_out/ord2_counters_cfg5.txt
