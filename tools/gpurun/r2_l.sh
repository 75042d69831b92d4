mkdir -p gpurun_out
run() { # name, extra env, extra args
GCNB_DP_SKIP_ALLREDUCE=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 40 --no-cpu-baseline $3 > gpurun_out/r2l_$1.json 2> gpurun_out/r2l_$1.err; echo "$1 rc=$?"
python - $1 <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2l_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('allreduce'), d['clocks'])
except Exception as e:
    print('parse error', e); print(open('gpurun_out/r2l_%s.err'%sys.argv[1]).read()[-1500:])
PY
}
run skip 1 ""
run peer 0 ""
run nccl 0 --nccl-allreduce
timeout 300 python bench.py --steps 400 --warmup 40 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single ms/step', round(d['ms_per_step'],4), d['clocks'])"
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 400 --warmup 40 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single gpu1 ms/step', round(d['ms_per_step'],4), d['clocks'])"
