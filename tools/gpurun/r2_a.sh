# round 2, first contact of the tcgen05 forward: parity suite, per-layer timings, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log; echo "tests rc=$?"
tail -5 gpurun_out/r2a_tests.log
WHICH="f1 f2 s1" timeout 300 python tools/time_layers.py 2>&1 | tail -5 | tee gpurun_out/r2a_layers.log
GCNB_UMMA=0 WHICH="f1 f2 s1" timeout 300 python tools/time_layers.py 2>&1 | tail -5 | tee gpurun_out/r2a_layers_old.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2a_bench.json
