#!/usr/bin/env python
"""Benchmark of the graph-convolution hot path (BASELINE.json metric: ChebyNet windows/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1): BASELINE config 2 -- ChebyNet K=5 training step (forward + backward + Adam),
2 conv layers F=[32,32] K=[5,5] p=[4,4] on the pinned 360-ROI graph coarsened 4 levels
(M=400->100->25), head 25->512->256->22, batch 512 windows of [360 ROI x 15 TR], fp32.
N>1: every rank runs the same per-GPU batch (weak scaling; N=8 is BASELINE config 4's global
batch 4096) and the filter/head gradients are averaged by one NCCL all-reduce per step.

One JSON line on stdout (rank 0).  ``value`` = windows/s with inputs resident in HBM (a ring of
distinct batches larger than 2x L2, so every step reads HBM); ``e2e`` = the same step fed from
pinned host memory with the H2D copy and a D2H read of the loss inside the timed region;
``roofline`` = algorithmic bytes / CUDA-event time of the dominant conv kernel against the
measured HBM peak; ``cpu_baseline`` = the NumPy oracle of the same step on the host cores.
``--impl reference`` times that CPU implementation alone (the reference's TF-1.x stack cannot be
installed).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 512
F, K, P, MFC = [32, 32], [5, 5], [4, 4], [512, 256, 22]
REG = 5e-4
METRIC = "ChebyNet K=5 training windows/sec (fwd+bwd+update, 2 conv layers + Graclus mpool, batch 512/GPU)"
WORKLOAD = "config2: ChebyNet K=5 train step, F=[32,32] K=[5,5] p=[4,4], M=400->100->25, head 25-512-256-22, B=512/GPU, 15-TR windows"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU side
def cpu_network(seed=7):
    from gcn_fmri_decoding_b200 import graclus, synth
    from oracle import layers_np as O

    A, gs, perm, L = synth.brain_graph(4)
    Ls = O.select_laplacians(L, P)
    rng = np.random.RandomState(seed)
    params, fin = [], 15
    for f, k, p in zip(F, K, P):
        params.append(dict(W=synth.truncated_normal(rng, (fin * k, f), 0.2), b=np.full(f, 0.2, np.float32), K=k, p=p))
        fin = f
    fcs, width = [], Ls[-1].shape[0] // P[-1]
    for m in MFC:
        fcs.append((synth.truncated_normal(rng, (width, m), 0.2), np.full(m, 0.2, np.float32)))
        width = m
    return perm, Ls, params, fcs


def cpu_step_rate(n_windows, reps, warm, cores, mode=None):
    """Seconds per oracle training step (fwd+loss+bwd; fp32 as the reference runs) on `cores` host threads.

    SciPy's CSR x dense product is single-threaded, so the step is parallelised over 32-window chunks by a thread
    pool; BLAS is pinned to ONE thread inside the pool (cores x cores oversubscription made the round-1 number 5x
    too slow).  mode "pool" = that; "blas" = no pool, BLAS on all cores; None = time both once, keep the faster.
    Returns (times, mode)."""
    from concurrent.futures import ThreadPoolExecutor

    from threadpoolctl import threadpool_limits

    from gcn_fmri_decoding_b200 import graclus, synth
    from oracle import layers_np as O

    perm, Ls, params, fcs = cpu_network()
    x = graclus.perm_data_3d(synth.bold_windows(n_windows, seed=2024), perm).astype(np.float32)
    lab = synth.labels(n_windows)
    chunks = [c for c in np.array_split(np.arange(n_windows), max(1, n_windows // 32)) if len(c)]

    def step(c):
        O.network_step(x[c], lab[c], Ls, params, fcs, REG, dtype=np.float32)

    def one_pool():
        with threadpool_limits(limits=1):
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(step, chunks))

    def one_blas():
        with threadpool_limits(limits=cores):
            step(np.arange(n_windows))

    def clock(fn):
        t = time.perf_counter()
        fn()
        return time.perf_counter() - t

    if mode is None:
        one_pool(), one_blas()
        mode = "pool" if clock(one_pool) <= clock(one_blas) else "blas"
    one = one_pool if mode == "pool" else one_blas
    for _ in range(warm):
        one()
    return [clock(one) for _ in range(reps)], mode


CPU_MODE_TEXT = {"pool": "thread pool over 32-window chunks, BLAS pinned to 1 thread per worker",
                 "blas": "one chunk, BLAS on all cores"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    probe, mode = cpu_step_rate(256, 1, 0, cores)
    rate = 256 / probe[0]
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(min(BATCH, max(32, (rate * budget) // 32 * 32)))
    times, mode = cpu_step_rate(n, args.steps, args.warmup, cores, mode)
    total = float(np.sum(times))
    value = n * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d windows per step" % n},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": "%d-window training step (fwd+loss+bwd, no optimiser update) of the NumPy/SciPy "
                                   "oracle, %d steps, %s" % (n, args.steps, CPU_MODE_TEXT[mode])},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU side
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        self.marks = []

    def mark(self):
        self.marks.append(self._lines())

    def _lines(self):
        self.file.flush()
        try:
            with open(self.file.name) as f:
                return sum(1 for _ in f)
        except OSError:
            return 0

    def stop(self, lo=None, hi=None):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        rows = []
        try:
            with open(self.file.name) as f:
                for ln in f:
                    parts = [p.strip() for p in ln.split(",")]
                    if len(parts) >= 8:
                        rows.append(parts)
            os.unlink(self.file.name)
        except OSError:
            pass
        sel = rows[lo:hi] if (lo is not None and hi is not None and hi > lo) else rows
        if not sel:
            sel = rows
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        sm = []
        for r in sel:
            try:
                sm.append(float(r[0]))
            except ValueError:
                pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in sel for i in range(4) if r[4 + i].lower().startswith("active")})
        try:
            mx = float(sel[0][1])
        except ValueError:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sel)}


def layer_bytes(B, M_read, M, Fin, Fout, Kk, p, nnz, nb, backward, need_dx):
    """ALGORITHMIC bytes of one conv-layer launch."""
    Mo = -(-M // p)
    csr = 8 * nnz + 4 * (M + 1)
    w = 4 * (Fin * Kk * Fout + nb)
    if not backward:
        return 4 * B * M_read * Fin + 4 * B * Mo * Fout + csr + w
    return 4 * B * M_read * Fin + 2 * 4 * B * Mo * Fout + (4 * B * M * Fin if need_dx else 0) + csr + 2 * w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 general (HBM) kernels, 2 fused kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cublas-fc", action="store_true", help="FC GEMMs through cuBLAS fp32 instead of the in-house 3xTF32 kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from gcn_fmri_decoding_b200 import _lib, ops, synth
    from gcn_fmri_decoding_b200.models import cgcnn
    from gcn_fmri_decoding_b200.train import FusedTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    A, gs, perm, L = synth.brain_graph(4)
    model = cgcnn(L=L, F=F, K=K, p=P, M=MFC, channel=15, device=dev, seed=7, regularization=REG, batch_size=BATCH,
                  perm=perm, n_input_vertices=360, algo=args.algo)
    # dropout keep-probability 0.5 on the FC layers, as the reference trains (model.py:169, models_gcn.py:145)
    trainer = FusedTrainer(model, use_cuda_graph=not args.no_graph, dropout=0.5, own_gemm=not args.cublas_fc)

    # ring of distinct resident batches: R x 11.06 MB of raw windows  (> 2 x 126 MB L2)
    R = 25
    host = synth.bold_windows(BATCH * 4, seed=2024 + rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring_x = [torch.as_tensor(host[(i % 4) * BATCH:(i % 4 + 1) * BATCH], device=dev)
              + 0.01 * torch.randn(BATCH, 360, 15, device=dev, generator=gen) for i in range(R)]
    ring_y = [torch.as_tensor(synth.labels(BATCH, seed=i + 100 * rank), device=dev) for i in range(R)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches of this library per step (counted on an eager step outside the timed region, state restored;
    # the CUDA graph replays exactly these launches).  world > 1: the eager step contains the all-reduce, every
    # rank runs it.
    snap = [t.clone() for t in (trainer.flat_p, trainer.flat_m, trainer.flat_v, trainer.state)]
    c0 = lib.gcnb_launch_count()
    trainer._step_impl(ring_x[0], ring_y[0])
    torch.cuda.synchronize()
    launches_per_step = int(lib.gcnb_launch_count() - c0)
    with torch.no_grad():
        for t, q in zip((trainer.flat_p, trainer.flat_m, trainer.flat_v, trainer.state), snap):
            t.copy_(q)

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        trainer.step(ring_x[i % R], ring_y[i % R])
    barrier()
    if sampler:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss, _ = trainer.step(ring_x[i % R], ring_y[i % R])
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = world * BATCH * args.steps / (ms_total * 1e-3)
    if sampler:
        sampler.mark()
    # keep the GPU under the same load for >= 1.5 s so that nvidia-smi (100 ms period) sees the clocks.
    # Every rank runs the same number of extra steps (the step contains a collective).
    if ms_total < 1500:
        extra = int(min(20000, np.ceil(1500.0 / max(ms_total / args.steps, 1e-3))))
        for i in range(extra):
            trainer.step(ring_x[i % R], ring_y[i % R])
            if i % 50 == 49:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        if sampler:
            sampler.marks[-1] = sampler._lines()
    final_loss = float(loss)
    assert np.isfinite(final_loss), "training diverged"

    # ---- end to end: pinned host ring -> H2D -> step -> D2H loss, all inside the timed region -------------
    pin_x = [torch.as_tensor(host[(i % 4) * BATCH:(i % 4 + 1) * BATCH]).pin_memory() for i in range(4)]
    pin_y = [torch.as_tensor(synth.labels(BATCH, seed=i)).pin_memory() for i in range(4)]
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    dbuf_x = [torch.empty(BATCH, 360, 15, device=dev) for _ in range(2)]
    dbuf_y = [torch.empty(BATCH, dtype=torch.long, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n, record):
        main_stream = torch.cuda.current_stream()
        for s in range(2):
            freed[s].record(main_stream)
        for i in range(n):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                dbuf_x[s].copy_(pin_x[i % 4], non_blocking=True)
                dbuf_y[s].copy_(pin_y[i % 4], non_blocking=True)
                ready[s].record(copy_stream)
            main_stream.wait_event(ready[s])
            l, _ = trainer.step(dbuf_x[s], dbuf_y[s])
            freed[s].record(main_stream)
            if record:
                loss_host[i].copy_(l, non_blocking=True)

    e2e_loop(max(3, args.warmup // 2), False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps, True)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * args.steps / (float(ms2) * 1e-3)
    h2d = BATCH * 360 * 15 * 4 + BATCH * 8
    assert np.all(np.isfinite(loss_host.numpy()))

    clocks = sampler.stop(*sampler.marks[-2:]) if sampler else None

    # ---- roofline of the conv kernels: CUDA events around repeated launches on the launching stream ---------
    roofline, kernels = None, []
    if rank == 0:
        peak, peak_src = measured_peaks()
        pl1, pl2 = model._plan(model.L[0]), model._plan(model.L[1])
        W1, W2, b1, b2 = model.conv_weights[0], model.conv_weights[1], model.conv_bias[0], model.conv_bias[1]
        mode = ops.BIAS_PER_FILTER
        n1 = 25
        with torch.no_grad():
            # the same launches the training step issues (train.FusedTrainer): forward keeps the Chebyshev basis,
            # backward = streamed dW GEMM from the basis (+ the adjoint recursion for dx in layer 2)
            f1 = lambda i: ops.cheb_fwd_mean(ring_x[i], model.perm, pl1.rowptr, pl1.col, pl1.val, W1, b1, 5, 4, mode, True,
                                             args.algo, True)
            ys1 = [f1(i) for i in range(n1)]
            n2 = 40
            y1s = [ys1[i % n1][0] + 0.0 * i for i in range(n2)]
            f2 = lambda i: ops.cheb_fwd_mean(y1s[i], None, pl2.rowptr, pl2.col, pl2.val, W2, b2, 5, 4, mode, True, args.algo,
                                             True)
            ys2 = [f2(i) for i in range(n2)]
            dm2 = [torch.randn_like(ys2[i][2]) for i in range(n2)]   # gradient of the mean over filters
            dy1 = [torch.randn_like(ys1[i][0]) for i in range(n1)]
            gW1, gb1 = torch.empty_like(W1), torch.empty(32, device=dev)
            gW2, gb2 = torch.empty_like(W2), torch.empty(32, device=dev)

            def timed(fn, n, reps=3):
                """GPU seconds per call: the n calls (one per ring slot) are captured into a CUDA graph and the
                graph replays are bracketed by CUDA events on the launching stream, so host launch overhead
                (tens of microseconds per Python op call) is not counted as kernel time."""
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for i in range(3):
                        fn(i % n)
                torch.cuda.current_stream().wait_stream(side)
                c0 = lib.gcnb_launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(n):
                        fn(i)
                per_call = (lib.gcnb_launch_count() - c0) / n
                g.replay()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for r in range(reps):
                    g.replay()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) * 1e-3 / (reps * n), per_call

            specs = [
                ("conv1 fwd: k_cheb_fwd_fused (gather+cheb K=5 15->32+b1relu+mpool4, M=400, keeps basis)", n1, f1,
                 layer_bytes(BATCH, 360, 400, 15, 32, 5, 4, pl1.nnz, 32, False, False)),
                ("conv2 fwd: k_cheb_fwd_fused (cheb K=5 32->32+b1relu+mpool4+mean, M=100, keeps basis)", n2, f2,
                 layer_bytes(BATCH, 100, 100, 32, 32, 5, 4, pl2.nnz, 32, False, False)),
                ("conv2 bwd: k_dw_from_stack + k_cheb_bwd_fused (dW,db from basis; dx by adjoint recursion)", n2,
                 lambda i: ops.cheb_bwd_into(y1s[i], None, ys2[i][0], ys2[i][1], dm2[i], True, *pl2.tensors(), W2, gW2, gb2, 5,
                                             4, mode, True, True, args.algo, ys2[i][3]),
                 layer_bytes(BATCH, 100, 100, 32, 32, 5, 4, pl2.nnz, 32, True, True)),
                ("conv1 bwd: k_dw_from_stack (dW,db from basis)", n1,
                 lambda i: ops.cheb_bwd_into(ring_x[i], model.perm, ys1[i][0], ys1[i][1], dy1[i], False, *pl1.tensors(), W1,
                                             gW1, gb1, 5, 4, mode, True, False, args.algo, ys1[i][3]),
                 layer_bytes(BATCH, 360, 400, 15, 32, 5, 4, pl1.nnz, 32, True, False)),
            ]
            traffic = {}
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.isfile(tpath):
                with open(tpath) as f:
                    traffic = json.load(f)
            for name, n, fn, nbytes in specs:
                sec, per_call = timed(fn, n)
                key = name.split(":")[0]
                kernels.append({"op": name, "launches_per_op": per_call, "us": sec * 1e6, "algorithmic_bytes": nbytes,
                                "achieved_gbs": nbytes / sec * 1e-9, "frac": nbytes / sec * 1e-9 / peak,
                                "traffic": traffic.get(key)})
        top = max(kernels, key=lambda k: k["us"])
        total_bytes = sum(k["algorithmic_bytes"] for k in kernels)
        total_us = sum(k["us"] for k in kernels)
        roofline = {"bound": "hbm", "kernel": top["op"], "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": top["frac"], "traffic": top["traffic"], "peak_source": peak_src,
                    "note": "algorithmic bytes per SURVEY 8d (no basis stack); the training path also writes/reads the "
                            "K-order basis (65.5 MB conv1, 32.8 MB conv2) -- see traffic and DESIGN.md section 3",
                    "conv_stack_fwd_bwd": {"algorithmic_bytes": total_bytes, "us": total_us,
                                           "achieved": total_bytes / total_us * 1e-3,
                                           "frac": total_bytes / total_us * 1e-3 / peak},
                    "kernels": kernels}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times, mode = cpu_step_rate(BATCH, 5, 2, cores)
        cpu = {"value": BATCH / float(np.median(times)), "unit": "windows/s", "cores": cores, "kind": "port",
               "sample": "512-window training step (fwd+loss+bwd, no optimiser update) of the NumPy/SciPy oracle, median "
                         "of 5 after 2 warm-ups, " + CPU_MODE_TEXT[mode]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                       "l2_policy": "inputs rotate through a ring of %d distinct batches (%.0f MB > 2x L2)" % (R, R * 11.06),
                       "cuda_graph": not args.no_graph, "algo": args.algo, "dropout_keep": 0.5,
                       "final_loss": final_loss},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "wall_s": wall},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Rank 0 alone runs the roofline / CPU legs; the others wait for it on the rendezvous store (a host-side
        # wait: no collective is left pending on the device).  The processes then leave without tearing the
        # communicator down: destroying a process group whose collectives live in captured CUDA graphs can hang.
        import datetime

        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("gcnb_bench_done", "1")
        else:
            store.wait(["gcnb_bench_done"], datetime.timedelta(minutes=30))
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
