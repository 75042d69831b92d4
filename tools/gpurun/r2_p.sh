mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2p_tests.log; echo "tests rc=$?"
tail -12 gpurun_out/r2p_tests.log
GCNB_IMAGE=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2p_img0.json 2> gpurun_out/r2p_img0.err; echo rc=$?
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2p_img1.json 2> gpurun_out/r2p_img1.err; echo rc=$?
python - <<'P'
import json
for n in ("img0","img1"):
    try:
        d=json.loads(open(f"gpurun_out/r2p_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(n, "fail", e); print(open(f"gpurun_out/r2p_{n}.err").read()[-1500:])
P
