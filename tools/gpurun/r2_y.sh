mkdir -p gpurun_out
for v in 8 2 1; do
GCNB_UMMA_NS_MAX=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2y_$v.json 2> gpurun_out/r2y_$v.err; echo rc=$?
python - $v <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2y_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("nsmax",sys.argv[1],d["ms_per_step"], d["value"], [(k["op"][:9], round(k["us"],1)) for k in d["roofline"]["kernels"]])
except Exception as e:
    print("fail", e); print(open("gpurun_out/r2y_%s.err"%sys.argv[1]).read()[-1500:])
P
done
