"""A few eager training steps of the config-2 network (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_fmri_decoding_b200 import synth
from gcn_fmri_decoding_b200.models import cgcnn
from gcn_fmri_decoding_b200.train import FusedTrainer
dev = torch.device("cuda:0")
A, gs, perm, L = synth.brain_graph(4)
m = cgcnn(L=L, F=[32, 32], K=[5, 5], p=[4, 4], M=[512, 256, 22], channel=15, device=dev, seed=7, regularization=5e-4,
          batch_size=512, perm=perm, n_input_vertices=360)
tr = FusedTrainer(m, use_cuda_graph=False, dropout=0.5, own_gemm=True, distributed=False)
x = torch.as_tensor(synth.bold_windows(512, seed=1), device=dev)
y = torch.as_tensor(synth.labels(512, seed=1), device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    tr.step(x, y)
torch.cuda.synchronize()
print("done")
