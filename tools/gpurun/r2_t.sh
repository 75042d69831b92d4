mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2t_tests.log; tail -4 gpurun_out/r2t_tests.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo rc=$?
python - <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
except Exception as e:
    print("fail", e); print(open("gpurun_out/r2t_bench.err").read()[-2000:])
P
