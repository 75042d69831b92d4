ncu --set full --clock-control none --import-source on -k regex:"k_cheb_fwd_fused|k_cheb_bwd_fused|k_dw_from_stack" -s 5 -c 5 -o gpurun_out/prof_step_r01 python tools/prof_step.py 3 > gpurun_out/prof_step.log 2>&1; tail -2 gpurun_out/prof_step.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 120 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
tail -1 gpurun_out/launch_bench.log | cut -c1-150
