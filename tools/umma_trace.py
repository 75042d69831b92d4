"""Debug (build with GCNB_NVCC_EXTRA=-DGCNB_TRACE): timeline of CTA 0 of k_cheb_fwd_umma for the conv1 / conv2 shapes."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gcn_fmri_decoding_b200 import _lib, ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan

dev = torch.device("cuda:0")
A, gs, perm, L = synth.brain_graph(4)
which = sys.argv[1] if len(sys.argv) > 1 else "f1"
if which == "f1":
    pl = GraphPlan(L[0], dev)
    x = torch.randn(512, 360, 15, device=dev)
    W = torch.randn(75, 32, device=dev) * .2
    pt = torch.as_tensor(perm, dtype=torch.int32, device=dev)
elif which == "f2":
    pl = GraphPlan(L[2], dev)
    x = torch.randn(512, 100, 32, device=dev)
    W = torch.randn(160, 32, device=dev) * .2
    pt = None
else:  # "c1": a middle layer of the config-1 network (M=372, 32->32, p=1, b2relu, B=128, inference: no basis kept)
    A1, gs1, perm1, L1 = synth.brain_graph(1)
    pl = GraphPlan(L1[0], dev)
    x = torch.randn(128, 372, 32, device=dev)
    W = torch.randn(160, 32, device=dev) * .2
    pt = None
if which in ("c1", "c1s"):
    b = torch.full((372, 32), .2, device=dev)
    print(_lib.describe_fwd(128, 372, pl.nnz, 32, 32, 5, 1))
    for _ in range(3):
        if which == "c1s":  # layers 2-6 of the config-1 network in one launch
            ops.cheb_stack_fwd(x, pl.rowptr, pl.col, pl.val, [W] * 5, [b] * 5, 5, 2, True)
        else:
            ops.cheb_fwd(x, pt, *pl.tensors(), W, b, 5, 1, 2, True, False, 0)
else:
    b = torch.full((32,), .2, device=dev)
    for _ in range(3):
        ops.cheb_fwd(x, pt, *pl.tensors(), W, b, 5, 4, 1, True, True, 0)
torch.cuda.synchronize()
lib = _lib.lib()
lib._handle if False else None
h = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * (4 * 512))()
rc = h.gcnb_debug_read_trace(buf)
t = np.array(buf[:]).reshape(4, 512)
t0 = t[t > 0].min()
names = ["sparse0", "items", "mma", "epi0"]
for r in range(4):
    if r == 1:
        v = t[1][t[1] > 0] - t0
        print("prologue", " ".join(str(int(a)) for a in t[1][:16][t[1][:16] > 0] - t0))
        w = t[1][64:64 + 20 * 16].reshape(20, 16)
        if (w > 0).any():
            base = w[w > 0].min()
            print("order n=2, per sparse warp: start | item0: loop, wait, finish | item1 ... | end, fenced, released (cycles from", int(base - t0), ")")
            for sw in range(20):
                print("  sw%2d" % sw, " ".join("%6d" % (int(a - base) if a > 0 else -1) for a in w[sw]))
        continue
    v = t[r][t[r] > 0] - t0
    print(names[r], len(v), " ".join(str(int(a)) for a in v[:64]))

for ti in range(2):
    w = t[3][64 + ti * 160:64 + ti * 160 + 160].reshape(20, 8)
    if (w > 0).any():
        base = w[w > 0].min()
        print("order 0 of tile %d, per sparse warp: start, staged | item0 item1 done | (5) end, fenced, released (cycles from %d)" % (ti, int(base - t0)))
        for sw in range(20):
            print("  o0 sw%2d" % sw, " ".join("%6d" % (int(a - base) if a > 0 else -1) for a in w[sw]))

w = t[0][256:256 + 24 * 8].reshape(24, 8)
if (w > 0).any():
    base = w[w > 0].min()
    print("first layer boundary, per warp (sparse 0-19, epilogue 20-23): start, loads issued, accumulators final, taps stored, drained, fenced, released (cycles from %d)" % int(base - t0))
    for i in range(24):
        print("  lb w%2d" % i, " ".join("%6d" % (int(a - base) if a > 0 else -1) for a in w[i][:7]))
