timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_f5.json 2> gpurun_out/bench_f5.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_f5.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['gpu_launches_per_step'], 'loss', d['config']['final_loss'])
PY
tail -3 gpurun_out/bench_f5.err
