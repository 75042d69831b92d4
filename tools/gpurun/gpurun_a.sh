ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
tail -1 gpurun_out/launch_bench.log | cut -c1-200
