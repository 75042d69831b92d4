# 8-GPU validation of the data-parallel step (bound-input graphs, peer-memory all-reduce), weak scaling
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2n8.json 2> gpurun_out/r2n8.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2n8.json').read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('allreduce'), d['config'].get('global_batch'))
except Exception as e:
    print('parse error', e); print(open('gpurun_out/r2n8.err').read()[-1500:])
PY
