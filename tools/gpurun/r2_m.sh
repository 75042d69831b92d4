mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -x -q 2>&1 | tail -8
run() {
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 40 --no-cpu-baseline $2 > gpurun_out/r2m_$1.json 2> gpurun_out/r2m_$1.err; echo "$1 rc=$?"
python - $1 <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2m_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('allreduce'))
except Exception as e:
    print('parse error', e); print(open('gpurun_out/r2m_%s.err'%sys.argv[1]).read()[-1500:])
PY
}
run peer ""
run nccl --nccl-allreduce
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
