# last check of the round: GPU tests, smoke(), evaluate() on the device, the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
python - <<'P' 2>&1 | tail -2
import numpy as np, torch
from gcn_fmri_decoding_b200 import synth, checkpoints
from gcn_fmri_decoding_b200.models import cgcnn
A, gs, perm, L = synth.brain_graph(4)
m = cgcnn(L=L, F=[32, 32], K=[5, 5], p=[4, 4], M=[512, 256, 22], channel=15, device="cuda", seed=1, batch_size=32, perm=perm, n_input_vertices=360)
x, y = synth.bold_windows(70, seed=2), synth.labels(70, seed=2)
f = checkpoints.save_checkpoint(m, "/tmp/ck/model", step=3)
print(m.evaluate(x, y, checkpoint=f)[0])
P
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), round(d['e2e']['value']))"
