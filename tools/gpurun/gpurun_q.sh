python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r01_default.json 2> gpurun_out/bench_r01_default.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_default.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches_per_step'], 'clocks', d['clocks'])
print('roofline top:', d['roofline']['kernel'][:60], round(d['roofline']['frac'],4), d['roofline']['traffic'])
for k in d['roofline']['kernels']: print('  ', k['op'][:70], k['launches_per_op'], round(k['us'],1), round(k['frac'],4))
print('conv total', d['roofline']['conv_stack_fwd_bwd'])
print('cpu', d['cpu_baseline'])
PY
tail -3 gpurun_out/bench_r01_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 | cut -c1-400
