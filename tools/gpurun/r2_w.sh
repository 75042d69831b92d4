mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2w_tests.log; tail -1 gpurun_out/r2w_tests.log
GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so timeout 300 python tools/umma_trace.py c1 2>&1 | grep -v "^  o0 sw\|^  sw" | tee gpurun_out/r2w_trace_c1.log
for c in 2 1; do
timeout 300 python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2w_bench$c.json 2> gpurun_out/r2w_bench$c.err; echo rc=$?
python - $c <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2w_bench%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("cfg",sys.argv[1],d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches_per_step"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
except Exception as e:
    print("fail", e); print(open("gpurun_out/r2w_bench%s.err"%sys.argv[1]).read()[-2000:])
P
done
