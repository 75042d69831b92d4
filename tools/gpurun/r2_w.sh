mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "random_shapes" 2>&1 | tail -30 > gpurun_out/r2w_tests.log; tail -12 gpurun_out/r2w_tests.log
