mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2o_tests.log; echo "tests rc=$?"
tail -4 gpurun_out/r2o_tests.log
GCNB_UMMA_ADJ=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2o_adj0.json 2> gpurun_out/r2o_adj0.err; echo rc=$?
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2o_adj1.json 2> gpurun_out/r2o_adj1.err; echo rc=$?
python - <<'P'
import json
for n in ("adj0","adj1"):
    try:
        d=json.loads(open(f"gpurun_out/r2o_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(n, "fail", e)
P
