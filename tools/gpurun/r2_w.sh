mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2w_tests.log; tail -1 gpurun_out/r2w_tests.log
for c in 1 2; do
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
