timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
ncu --set full --clock-control none --import-source on -k regex:k_cheb_bwd_fused -s 6 -c 2 -o gpurun_out/prof_bwd1 python tools/prof_layers.py bwd 4 2 > gpurun_out/prof_bwd1.log 2>&1; tail -2 gpurun_out/prof_bwd1.log
