"""Run the conv-layer ops of the config-2 network a few times (target for ncu)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_fmri_decoding_b200 import ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
algo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
A, gs, perm, L = synth.brain_graph(4)
B = 512
pl1, pl2 = GraphPlan(L[0], dev), GraphPlan(L[2], dev)
permt = torch.as_tensor(perm, dtype=torch.int32, device=dev)
xraw = torch.randn(B, 360, 15, device=dev)
W1 = torch.randn(75, 32, device=dev) * 0.2
W2 = torch.randn(160, 32, device=dev) * 0.2
b = torch.full((32,), 0.2, device=dev)
mode = ops.BIAS_PER_FILTER
for _ in range(reps):
    y1, a1 = ops.cheb_fwd(xraw, permt, *pl1.tensors(), W1, b, 5, 4, mode, True, True, algo)
    y2, a2 = ops.cheb_fwd(y1, None, *pl2.tensors(), W2, b, 5, 4, mode, True, True, algo)
    if which in ("all", "bwd"):
        dy2 = torch.randn_like(y2)
        dx2, dW2, db2 = torch.ops.gcn_b200.cheb_bwd(y1, None, y2, a2, dy2, *pl2.tensors(), W2, 5, 4, mode, True, True, algo)
        dx1, dW1, db1 = torch.ops.gcn_b200.cheb_bwd(xraw, permt, y1, a1, dx2, *pl1.tensors(), W1, 5, 4, mode, True, False, algo)
torch.cuda.synchronize()
print("done")
