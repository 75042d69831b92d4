mkdir -p gpurun_out
WHICH="f1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cheb_fwd_umma -s 2 -c 2 -o gpurun_out/r2b_umma_f1 -f python tools/time_layers.py > gpurun_out/r2b_ncu1.log 2>&1; echo "ncu f1 rc=$?"
WHICH="f2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cheb_fwd_umma -s 2 -c 1 -o gpurun_out/r2b_umma_f2 -f python tools/time_layers.py > gpurun_out/r2b_ncu2.log 2>&1; echo "ncu f2 rc=$?"
tail -3 gpurun_out/r2b_ncu1.log
