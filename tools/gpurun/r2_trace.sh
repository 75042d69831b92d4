mkdir -p gpurun_out
timeout 300 python tools/umma_trace.py f1 > gpurun_out/r2_trace_f1.log 2>&1; echo rc=$?
timeout 300 python tools/umma_trace.py f2 > gpurun_out/r2_trace_f2.log 2>&1; echo rc=$?
cat gpurun_out/r2_trace_f1.log gpurun_out/r2_trace_f2.log
