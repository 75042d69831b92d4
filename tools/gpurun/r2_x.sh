mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/r02f_step -f python tools/prof_step.py 5 > gpurun_out/r02f_step_ncu.log 2>&1; echo "step ncu rc=$?"; grep -i "error" gpurun_out/r02f_step_ncu.log | head -3
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 27 -c 18 --csv --log-file gpurun_out/r02f_launches.csv python tools/prof_step.py 5 > gpurun_out/r02f_launches.log 2>&1; echo "launch list rc=$?"
