mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_rollout_kernel -s 2 -c 1 -f -o gpurun_out/e2_rollout_cfg5 python tools/profile_solve.py cfg5 > gpurun_out/e2_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_accept_kernel -s 2 -c 1 -f -o gpurun_out/e2_accept_cfg5 python tools/profile_solve.py cfg5 > gpurun_out/e2_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_sweep_kernel -s 2 -c 1 -f -o gpurun_out/e2_sweep_cfg5 python tools/profile_solve.py cfg5 > gpurun_out/e2_full3.log 2>&1
ls -la gpurun_out/e2_*.ncu-rep
