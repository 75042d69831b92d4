mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2y_tests.log; tail -1 gpurun_out/r2y_tests.log
for v in 0 1; do
GCNB_FORK_PARTIALS=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2y_$v.json 2> gpurun_out/r2y_$v.err; echo rc=$?
python - $v <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2y_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("fork",sys.argv[1],d["ms_per_step"], d["value"], d["e2e"]["value"], [(k["op"][:9], round(k["us"],1)) for k in d["roofline"]["kernels"]])
except Exception as e:
    print("fail", e); print(open("gpurun_out/r2y_%s.err"%sys.argv[1]).read()[-1500:])
P
done
