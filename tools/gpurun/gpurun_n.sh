timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in "" "--cublas-fc"; do python bench.py --steps 100 --warmup 10 --no-cpu-baseline $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])"; done
