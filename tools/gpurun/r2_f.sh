mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2f_tests.log; echo "tests rc=$?"
tail -5 gpurun_out/r2f_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches_per_step'], 'clocks', d['clocks'])
for k in d['roofline']['kernels']: print('  ', k['op'][:70], k['launches_per_op'], round(k['us'],1), round(k['frac'],4))
print('cpu', d['cpu_baseline'])
PY
tail -3 gpurun_out/r2f_bench.err
