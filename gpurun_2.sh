set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f1.json 2> gpurun_out/bench_f1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_f1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
for k in d['roofline']['kernels']: print(k['op'], k['launches_per_op'], round(k['us'],1), round(k['frac'],4))
PY
tail -5 gpurun_out/bench_f1.err
