mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pdl in 1 0; do
GCNB_PDL=$pdl timeout 600 python bench.py --no-cpu-baseline --steps 400 --warmup 40 > gpurun_out/r2n_pdl$pdl.json 2> gpurun_out/r2n_pdl$pdl.err; echo "pdl=$pdl rc=$?"
python - $pdl <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2n_pdl%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), [round(k['us'],1) for k in d['roofline']['kernels']])
except Exception as e:
    print('parse error', e); print(open('gpurun_out/r2n_pdl%s.err'%sys.argv[1]).read()[-1500:])
PY
done
