bash gcn_fmri_decoding_b200/csrc/build.sh > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_head_step -s 3 -c 1 -o gpurun_out/r2h_head -f python tools/time_head.py > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
