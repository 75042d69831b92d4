mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "layer_stack or uses_the_layer_stack" 2>&1 | tail -12 > gpurun_out/r2w_stack.log; echo "stack tests rc=$?"; tail -6 gpurun_out/r2w_stack.log
timeout 100 python bench.py --config 1 --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/r2w_bench1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', d['ms_per_step'], round(d['value']), d['gpu_launches_per_step'], [(k['op'][:12], round(k['us'],1)) for k in d['roofline']['kernels']])"
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_layer or fused_trainer_with_dropout or model_logits" 2>&1 | tail -1
