mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2t_tests.log; tail -6 gpurun_out/r2t_tests.log
for v in 0 1; do
GCNB_PDL=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2t_bench$v.json 2> gpurun_out/r2t_bench$v.err; echo rc=$?
python - $v <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2t_bench%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("pdl",sys.argv[1],d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches_per_step"], [(k["op"][:14], round(k["us"],1)) for k in d["roofline"]["kernels"]])
except Exception as e:
    print("fail", e); print(open("gpurun_out/r2t_bench%s.err"%sys.argv[1]).read()[-2000:])
P
done
