import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_fmri_decoding_b200 import ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan
dev = torch.device('cuda:0')
A, gs, perm, L = synth.brain_graph(4)
pl1 = GraphPlan(L[0], dev)
permt = torch.as_tensor(perm, dtype=torch.int32, device=dev)
x = torch.randn(512, 360, 15, device=dev)
W1 = torch.randn(75, 32, device=dev) * .2
b = torch.full((32,), .2, device=dev)
for _ in range(3):
    y, am, hm, st = ops.cheb_fwd_mean(x, permt, pl1.rowptr, pl1.col, pl1.val, W1, b, 5, 4, 1, True, 2, True)
    dy = torch.randn_like(y)
    gW, gb = torch.empty_like(W1), torch.empty(32, device=dev)
    ops.cheb_bwd_into(x, permt, y, am, dy, False, *pl1.tensors(), W1, gW, gb, 5, 4, 1, True, False, 2, st)
torch.cuda.synchronize()
