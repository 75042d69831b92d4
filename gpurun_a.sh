ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
tail -2 gpurun_out/launch_bench.log | cut -c1-300
