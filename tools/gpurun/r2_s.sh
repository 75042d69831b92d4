mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2s_tests.log; tail -3 gpurun_out/r2s_tests.log
for w in f1 f2; do
echo "== $w"
GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so timeout 300 python tools/umma_trace.py $w 2>&1 | grep -v "  sw1[0-4]\|  sw [2-9]" | tee gpurun_out/r2s_trace_$w.log
done
for st in 0 1; do
GCNB_STAGE=$st timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2s_bench$st.json 2> gpurun_out/r2s_bench$st.err; echo rc=$?
python - $st <<'P'
import json,sys
d=json.loads(open("gpurun_out/r2s_bench%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print("stage",sys.argv[1],d["ms_per_step"], d["value"], d["e2e"]["value"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
P
done
