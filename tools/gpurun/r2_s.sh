mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2s_tests.log; tail -1 gpurun_out/r2s_tests.log
for w in f1 f2; do
echo "== $w"
GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so timeout 300 python tools/umma_trace.py $w 2>&1 | grep -v "  sw1[0-4]\|  sw [2-9]\|o0 sw" | tee gpurun_out/r2s_trace_$w.log
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo rc=$?
python - <<'P'
import json,sys
d=json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
P
