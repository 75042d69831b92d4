set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --steps 30 --warmup 5 --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; tail -c 3000 gpurun_out/bench_nograph.json; tail -5 gpurun_out/bench_nograph.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -c 1500 gpurun_out/bench_graph.json; tail -5 gpurun_out/bench_graph.err
