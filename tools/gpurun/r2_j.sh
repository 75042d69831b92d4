mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_tests.log
for c in 2; do
  timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2j_bench_$c.json 2> gpurun_out/r2j_bench_$c.err; echo "config $c rc=$?"
  python - $c <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2j_bench_%s.json'%c).read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches_per_step'], 'roof', round(d['roofline']['frac'],4))
    for k in d['roofline'].get('kernels',[]): print('    ', k['op'][:80], round(k['us'],1))
except Exception as e:
    print(' parse error', e); print(open('gpurun_out/r2j_bench_%s.err'%c).read()[-600:])
PY
done
