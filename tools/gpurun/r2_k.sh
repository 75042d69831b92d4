mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -x -q 2>&1 | tail -15
for peer in 1 0; do
GCNB_PEER=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 $( [ $peer = 0 ] && echo --nccl-allreduce ) > gpurun_out/r2k_bench2_peer$peer.json 2> gpurun_out/r2k_bench2_peer$peer.err; echo "2gpu peer=$peer rc=$?"
python - $peer <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2k_bench2_peer%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('allreduce'))
except Exception as e:
    print('parse error', e); print(open('gpurun_out/r2k_bench2_peer%s.err'%sys.argv[1]).read()[-1500:])
PY
done
