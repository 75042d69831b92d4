mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_tests.log; echo "tests rc=$?"
tail -4 gpurun_out/r2c_tests.log
WHICH="f1 f2 s1" timeout 300 python tools/time_layers.py 2>&1 | tail -3 | tee gpurun_out/r2c_layers.log
WHICH="f1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cheb_fwd_umma -s 2 -c 1 -o gpurun_out/r2c_umma_f1 -f python tools/time_layers.py > gpurun_out/r2c_ncu1.log 2>&1; echo "ncu f1 rc=$?"
WHICH="f2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cheb_fwd_umma -s 2 -c 1 -o gpurun_out/r2c_umma_f2 -f python tools/time_layers.py > gpurun_out/r2c_ncu2.log 2>&1; echo "ncu f2 rc=$?"
