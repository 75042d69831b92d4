mkdir -p gpurun_out
timeout 60 python bench.py --config 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2w_bench1.json 2> gpurun_out/r2w_bench1.err; echo rc=$?
tail -c 600 gpurun_out/r2w_bench1.json | head -c 600; tail -3 gpurun_out/r2w_bench1.err
