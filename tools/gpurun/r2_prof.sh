# round-2 profile captures: launch list of one step, ncu --set full of the step's kernels, bench lines for every config
mkdir -p gpurun_out
# 1. launch list: every launch of 2 eager steps after 3 warm-up steps (cold caches, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 27 -c 18 --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py 5 > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
# 2. full sets for the kernels of one step
timeout 900 ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/r02_step -f python tools/prof_step.py 5 > gpurun_out/r02_step_ncu.log 2>&1; echo "step ncu rc=$?"
# 3. bench lines
for c in 2 3a 3a1 3b 1 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r02_bench_cfg$c.json 2> gpurun_out/r02_bench_cfg$c.err; echo "config $c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
