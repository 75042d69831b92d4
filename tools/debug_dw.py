"""Debug helper: general-path dW/dx/db error pattern against a torch fp64 reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gcn_fmri_decoding_b200 import ops, synth, _lib, graphs
from gcn_fmri_decoding_b200.plan import GraphPlan

dev = torch.device("cuda:0")
M = int(os.environ.get("M", 3000)); B = int(os.environ.get("B", 3)); K = int(os.environ.get("K", 25))
Fin = int(os.environ.get("FIN", 15)); Fout = int(os.environ.get("FOUT", 32))
L = synth.fibonacci_sphere_graph(M, 6)
pl = GraphPlan(L, dev)
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.randn(B, M, Fin, generator=g).to(dev)
W = (torch.randn(Fin * K, Fout, generator=g) * 0.05).to(dev)
bias = torch.full((Fout,), 0.2, device=dev)
dy = torch.randn(B, M, Fout, generator=g).to(dev)
if os.environ.get("DIRTY"):
    junk = torch.full((1 << 28,), float(os.environ["DIRTY"]), device=dev)
    del junk
if os.environ.get("TESTDATA"):
    rng = np.random.RandomState(9)
    x = torch.as_tensor(rng.randn(B, M, Fin).astype(np.float32), device=dev)
    W = torch.as_tensor((rng.randn(Fin * K, Fout) * 0.05).astype(np.float32), device=dev)
    dy = torch.as_tensor(rng.randn(B, M, Fout).astype(np.float32), device=dev)
with torch.no_grad():
    y, am = ops.cheb_fwd(x, None, *pl.tensors(), W, bias, K, 1, ops.BIAS_PER_FILTER, True, True, _lib.ALGO_GENERAL)
    dx, dW, db = torch.ops.gcn_b200.cheb_bwd(x, None, y, am, dy, *pl.tensors(), W, K, 1, ops.BIAS_PER_FILTER, True, True, _lib.ALGO_GENERAL)
# fp64 reference
Lt = graphs.rescale_L(L, 2).tocoo()
Ld = torch.sparse_coo_tensor(np.vstack([Lt.row, Lt.col]), Lt.data.astype(np.float64), Lt.shape).to(dev)
X0 = x.double().permute(1, 0, 2).reshape(M, B * Fin)
Xs = [X0]
if K > 1: Xs.append(torch.sparse.mm(Ld, X0))
for k in range(2, K): Xs.append(2 * torch.sparse.mm(Ld, Xs[-1]) - Xs[-2])
St = torch.stack(Xs).reshape(K, M, B, Fin)           # [K, M, B, Fin]
Wr = W.double().reshape(Fin, K, Fout)
z = torch.einsum("kmbf,fko->bmo", St, Wr) + 0.2
yr = z.clamp(min=0)
dZ = dy.double() * (z > 0)
dWr = torch.einsum("kmbf,bmo->fko", St, dZ).reshape(Fin * K, Fout)
print("y err", float((y - yr).abs().max() / yr.abs().max()))
e = (dW.double() - dWr).abs().reshape(Fin, K, Fout)
print("dW err", float(e.max() / dWr.abs().max()), "max|dW|", float(dWr.abs().max()))
print("err by k:", [round(float(v), 4) for v in e.amax((0, 2))])
print("err by f:", [round(float(v), 4) for v in e.amax((1, 2))])
print("err by o:", [round(float(v), 4) for v in e.amax((0, 1))])
print("db err", float((db.double() - dZ.sum((0, 1))).abs().max() / dZ.sum((0, 1)).abs().max()))
