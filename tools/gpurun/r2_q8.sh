# 8-GPU validation of the data-parallel step: NCCL test, bench at N=8 (peer-memory all-reduce and plain NCCL), strong mode
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -x -q 2>&1 | tail -4
run() {  # name, nproc, extra flags
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 200 --warmup 20 --no-cpu-baseline $3 > gpurun_out/r2q_$1.json 2> gpurun_out/r2q_$1.err; echo "$1 rc=$?"
python - $1 <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2q_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('allreduce'), d['config'].get('global_batch'))
except Exception as e:
    print('parse error', e); print(open('gpurun_out/r2q_%s.err'%sys.argv[1]).read()[-1500:])
PY
}
run n8_peer 8 ""
run n8_nccl 8 --nccl-allreduce
run n4_peer 4 ""
run n8_strong 8 --strong
