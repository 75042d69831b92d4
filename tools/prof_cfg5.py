"""One forward and one backward of BASELINE config 5 (vertex-level ChebyNet, K=25, 32 492 vertices, B=64).

    python tools/prof_cfg5.py            # prints CUDA-event times (ms) of fwd and bwd
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/cfg5_launches.csv python tools/prof_cfg5.py once
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gcn_fmri_decoding_b200 import ops, synth
from gcn_fmri_decoding_b200.plan import GraphPlan

once = len(sys.argv) > 1 and sys.argv[1] == "once"
M = int(os.environ.get("CFG5_M", 32492))
B = int(os.environ.get("CFG5_B", 64))
K = int(os.environ.get("CFG5_K", 25))
dev = torch.device("cuda:0")
Ls = synth.fibonacci_sphere_graph(M, 6)
pl = GraphPlan(Ls, dev)
x = torch.randn(B, M, 15, device=dev)
W = torch.randn(15 * K, 32, device=dev) * 0.05
bias = torch.full((32,), 0.2, device=dev)


def fwd():
    return ops.cheb_fwd(x, None, *pl.tensors(), W, bias, K, 1, ops.BIAS_PER_FILTER, True, True, 0)


with torch.no_grad():
    y, am = fwd()
    dy = torch.randn_like(y)

    def bwd():
        return torch.ops.gcn_b200.cheb_bwd(x, None, y, am, dy, *pl.tensors(), W, K, 1, ops.BIAS_PER_FILTER, True, True, 0)

    bwd()
    gW, gb = torch.empty_like(W), torch.empty(32, device=dev)

    def fwd_keep():
        return ops.cheb_fwd_mean(x, None, pl.rowptr, pl.col, pl.val, W, bias, K, 1, ops.BIAS_PER_FILTER, True, 0, True)

    stack = fwd_keep()[3]

    def bwd_saved():
        return ops.cheb_bwd_into(x, None, y, am, dy, False, *pl.tensors(), W, gW, gb, K, 1, ops.BIAS_PER_FILTER, True, True,
                                 0, stack)

    bwd_saved()
    if not once:
        for name, fn in (("fwd", fwd), ("bwd", bwd), ("fwd_keep_basis", fwd_keep), ("bwd_saved_basis", bwd_saved)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            print("%s %.3f ms  (%d windows -> %.0f windows/s)" % (name, a.elapsed_time(b) / 5, B, B / (a.elapsed_time(b) / 5e3)))
