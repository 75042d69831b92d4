"""GPU time of the four conv-layer launches of the config-2 step, host overhead excluded (CUDA-graph replay)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_fmri_decoding_b200 import ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan

algo = int(os.environ.get("ALGO", "2"))
dev = torch.device('cuda:0')
A, gs, perm, L = synth.brain_graph(4)
pl1, pl2 = GraphPlan(L[0], dev), GraphPlan(L[2], dev)
permt = torch.as_tensor(perm, dtype=torch.int32, device=dev)
n1, n2 = 25, 40
xr = [torch.randn(512, 360, 15, device=dev) for _ in range(n1)]
W1 = torch.randn(75, 32, device=dev) * .2
W2 = torch.randn(160, 32, device=dev) * .2
b = torch.full((32,), .2, device=dev)
y1 = [ops.cheb_fwd(xr[i], permt, *pl1.tensors(), W1, b, 5, 4, 1, True, True, algo) for i in range(n1)]
xs = [y1[i % n1][0] + 0.01 * i for i in range(n2)]
y2 = [ops.cheb_fwd(xs[i], None, *pl2.tensors(), W2, b, 5, 4, 1, True, True, algo) for i in range(n2)]
dy2 = [torch.randn_like(y2[i][0]) for i in range(n2)]
dy1 = [torch.randn_like(y1[i][0]) for i in range(n1)]


def timeit(f, n):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            f(i % n)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            f(i)
    g.replay()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for r in range(3):
        g.replay()
    c.record()
    torch.cuda.synchronize()
    return a.elapsed_time(c) * 1e3 / (3 * n)


which = os.environ.get("WHICH", "f1 f2 b2 b1").split()
out = []
if "f1" in which:
    out.append("L1 fwd %.1f" % timeit(lambda i: ops.cheb_fwd(xr[i], permt, *pl1.tensors(), W1, b, 5, 4, 1, True, True, algo), n1))
if "f2" in which:
    out.append("L2 fwd %.1f" % timeit(lambda i: ops.cheb_fwd(xs[i], None, *pl2.tensors(), W2, b, 5, 4, 1, True, True, algo), n2))
if "b2" in which:
    out.append("L2 bwd %.1f" % timeit(lambda i: torch.ops.gcn_b200.cheb_bwd(xs[i], None, y2[i][0], y2[i][1], dy2[i], *pl2.tensors(), W2, 5, 4, 1, True, True, algo), n2))
if "b1" in which:
    out.append("L1 bwd %.1f" % timeit(lambda i: torch.ops.gcn_b200.cheb_bwd(xr[i], permt, y1[i][0], y1[i][1], dy1[i], *pl1.tensors(), W1, 5, 4, 1, True, False, algo), n1))
if "s1" in which:
    st = [ops.cheb_fwd_mean(xr[i], permt, pl1.rowptr, pl1.col, pl1.val, W1, b, 5, 4, 1, True, algo, True) for i in range(4)]
    out.append("L1 fwd+stack %.1f" % timeit(lambda i: ops.cheb_fwd_mean(xr[i], permt, pl1.rowptr, pl1.col, pl1.val, W1, b, 5, 4, 1, True, algo, True), n1))
    gW, gb = torch.empty_like(W1), torch.empty(32, device=dev)
    out.append("L1 bwd(stack) %.1f" % timeit(lambda i: ops.cheb_bwd_into(xr[i], permt, y1[i][0], y1[i][1], dy1[i], False, *pl1.tensors(), W1, gW, gb, 5, 4, 1, True, False, algo, st[i % 4][3]), n1))
print(os.environ.get("TAG", ""), "us:", "  ".join(out))
