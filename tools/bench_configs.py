"""Timing of the other BASELINE.json configurations (1, 3a, 3b, 5) on one B200; writes one JSON object per config.

    python tools/bench_configs.py [out.json]

config 1: ChebyNet K=5 predict, 6 conv layers on M=372 (levels=1), b2relu, head 372-512-256-22, B=128 (forward only;
          CPU side = the NumPy oracle on the host cores)
config 3a: K=2 ("chebyshev2" entry) and K=1 ("firstorder") training step at the config-2 shape
config 3b: spectral (fourier) training step at the config-2 shape
config 5: vertex-level ChebyNet K=25 on a 32 492-vertex Fibonacci-sphere kNN graph, 15->32, B=64, forward and
          forward+backward (general HBM-resident path)
GPU times are CUDA-event times of CUDA-graph replays (host launch overhead excluded), inputs rotate through rings.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gcn_fmri_decoding_b200 import graclus, ops, synth
from gcn_fmri_decoding_b200.models import cgcnn
from gcn_fmri_decoding_b200.plan import GraphPlan
from gcn_fmri_decoding_b200.train import FusedTrainer

dev = torch.device("cuda:0")
out = {}


def graph_time(fn, n, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(2):
            fn(i % n)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / (reps * n)


def event_time(fn, reps=3):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


# ---------------------------------------------------------------- config 1
A, gs, perm, L = synth.brain_graph(1)
B = 128
model = cgcnn(L=L, F=[32] * 6, K=[5] * 6, p=[1] * 6, M=[512, 256, 22], filter="chebyshev5", brelu="b2relu", channel=15,
              device=dev, seed=7, batch_size=B, perm=perm, n_input_vertices=360)
ring = [torch.as_tensor(synth.bold_windows(B, seed=30 + i), device=dev) for i in range(16)]
with torch.no_grad():
    t = graph_time(lambda i: model(ring[i]), 16)
rec = {"workload": "ChebyNet K=5 predict, 6 conv layers M=372 b2relu + head, B=128, forward", "gpu_s_per_batch": t,
       "gpu_windows_per_s": B / t}
try:
    from concurrent.futures import ThreadPoolExecutor

    from oracle import layers_np as O

    sd = model.state_dict_tf()
    params = [dict(W=sd["conv%d/weights" % (i + 1)], b=sd["conv%d/bias" % (i + 1)].reshape(-1, 32), K=5, p=1) for i in range(6)]
    fcs = [(sd[s + "/weights"], sd[s + "/bias"]) for s in ("fc1", "fc2", "logits")]
    xp = graclus.perm_data_3d(synth.bold_windows(B, seed=30), perm).astype(np.float32)
    chunks = np.array_split(np.arange(B), 8)
    cores = os.cpu_count() or 1

    def one():
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda c: O.head(O.conv_stack(xp[c], model.L, params, brelu="b2relu", dtype=np.float32), fcs, np.float32), chunks))

    one()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    rec.update(cpu_windows_per_s=B / float(np.median(ts)), cpu_cores=cores, cpu_kind="port (NumPy/SciPy oracle)")
except Exception as e:  # the CPU leg is informational
    rec["cpu_error"] = repr(e)
out["config1_predict"] = rec
del model, ring

# ---------------------------------------------------------------- configs 3a / 3b (training step at the config-2 shape)
A, gs, perm, L = synth.brain_graph(4)
B = 512
xs = [torch.as_tensor(synth.bold_windows(B, seed=40 + i), device=dev) for i in range(8)]
ys = [torch.as_tensor(synth.labels(B, seed=40 + i), device=dev) for i in range(8)]
for name, filt, Ks in (("config3a_K2_chebyshev2", "chebyshev2", [2, 2]), ("config3a_K1_firstorder", "chebyshev5", [1, 1]),
                       ("config2_K5_chebyshev5", "chebyshev5", [5, 5]), ("config3b_spectral", "fourier", [0, 0])):
    model = cgcnn(L=L, F=[32, 32], K=Ks, p=[4, 4], M=[512, 256, 22], filter=filt, brelu="b1relu", channel=15, device=dev,
                  seed=7, regularization=5e-4, batch_size=B, perm=perm, n_input_vertices=360)
    tr = FusedTrainer(model, use_cuda_graph=True, dropout=0.5, distributed=False)
    for i in range(5):
        tr.step(xs[i % 8], ys[i % 8])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 50
    for i in range(n):
        loss, _ = tr.step(xs[i % 8], ys[i % 8])
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3 / n
    out[name] = {"workload": "%s K=%s training step (fwd+bwd+Adam), B=512" % (filt, Ks), "gpu_s_per_step": t,
                 "gpu_windows_per_s": B / t, "final_loss": float(loss)}
    del model, tr
del xs, ys
torch.cuda.empty_cache()

# ---------------------------------------------------------------- config 5
Ls = synth.fibonacci_sphere_graph(32492, 6)
pl = GraphPlan(Ls, dev)
B, K = 64, 25
x = torch.randn(B, 32492, 15, device=dev)
W = torch.randn(15 * K, 32, device=dev) * 0.05
bias = torch.full((32,), 0.2, device=dev)
with torch.no_grad():
    tf = event_time(lambda: ops.cheb_fwd(x, None, *pl.tensors(), W, bias, K, 1, ops.BIAS_PER_FILTER, True, False, 0))
    y, am = ops.cheb_fwd(x, None, *pl.tensors(), W, bias, K, 1, ops.BIAS_PER_FILTER, True, True, 0)
    dy = torch.randn_like(y)
    tb = event_time(lambda: torch.ops.gcn_b200.cheb_bwd(x, None, y, am, dy, *pl.tensors(), W, K, 1, ops.BIAS_PER_FILTER, True,
                                                        True, 0))
ideal_fwd = 4 * B * 32492 * (15 + 32) + 8 * pl.nnz
out["config5_vertex_level"] = {"workload": "ChebyNet K=25, M=32492 (nnz %d), 15->32, B=64" % pl.nnz, "fwd_s": tf,
                               "fwd_windows_per_s": B / tf, "fwd_bwd_s": tf + tb, "fwd_bwd_windows_per_s": B / (tf + tb),
                               "fwd_algorithmic_bytes": ideal_fwd, "fwd_achieved_gbs_vs_ideal_bytes": ideal_fwd / tf * 1e-9}
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
