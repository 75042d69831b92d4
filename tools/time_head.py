"""GPU time of gcnb_head_step_f32 at the config-2 sizes (B=512, 25-512-256-22), CUDA-graph replay of 20 calls."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_fmri_decoding_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.lib()
B, n0, n1, n2, nc = 512, 25, 512, 256, 22
a0 = torch.rand(B, n0, device=dev)
lab = torch.randint(0, nc, (B,), device=dev)
W = [torch.randn(a, b, device=dev) * 0.2 for a, b in ((n0, n1), (n1, n2), (n2, nc))]
bs = [torch.full((b,), 0.1, device=dev) for b in (n1, n2, nc)]
gW = [torch.empty_like(w) for w in W]
gb = [torch.empty_like(b) for b in bs]
logits = torch.empty(B, nc, device=dev)
loss = torch.zeros((), device=dev)
d0 = torch.empty(B, n0, device=dev)
state = torch.tensor([1.0, 1.0, 0.0, 0.0], device=dev)
ws = torch.empty(lib.gcnb_head_step_workspace_bytes(B, n0, n1, n2, nc), dtype=torch.uint8, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())


def call():
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.gcnb_head_step_f32(vp(a0), vp(lab), vp(W[0]), vp(bs[0]), vp(W[1]), vp(bs[1]), vp(W[2]), vp(bs[2]), vp(logits),
                                vp(loss), vp(gW[0]), vp(gb[0]), vp(gW[1]), vp(gb[1]), vp(gW[2]), vp(gb[2]), vp(d0), B, n0, n1,
                                n2, nc, 0.5, 1, 2, vp(state), 1e-3, 0.9, 0.999, 1, vp(ws), ws.numel(), st)
    _lib.check(rc, "head")


for _ in range(3):
    call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    call()
b.record()
torch.cuda.synchronize()
print("eager: %.1f us per call" % (a.elapsed_time(b) * 1e3 / 20))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    call()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        call()
g.replay()
torch.cuda.synchronize()
a.record()
for _ in range(3):
    g.replay()
b.record()
torch.cuda.synchronize()
print("graph: %.1f us per call" % (a.elapsed_time(b) * 1e3 / 60))

if os.environ.get("GCNB_HEAD_TRACE"):
    h = C.CDLL(_lib.LIB_PATH)
    buf = (C.c_longlong * 32)()
    h.gcnb_debug_read_head_trace(buf)
    v = [int(x) for x in buf[:14]]
    names = ["F1", "sync", "F2", "sync", "F3", "sync", "B1", "sync", "B2", "sync", "B3", "sync", "B4"]
    print("head phases (cycles):", " ".join("%s=%d" % (n, v[i + 1] - v[i]) for i, n in enumerate(names)))
    import numpy as np
    big = (C.c_longlong * (160 * 16))()
    h.gcnb_debug_read_head_cta(big)
    a = np.array(big[:]).reshape(160, 16)[:148]
    for i, n in enumerate(names):
        if n == "sync":
            continue
        d = a[:, i + 1] - a[:, i]
        order = np.argsort(-d)
        print("%s per CTA: median %d max %d; slowest CTAs %s" % (n, int(np.median(d)), int(d.max()),
              " ".join("%d:%d" % (int(b), int(d[b])) for b in order[:8])))
