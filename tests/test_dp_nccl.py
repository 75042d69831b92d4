"""Data-parallel training on real GPUs: two ranks over NCCL (skipped with fewer than two devices).

`FusedTrainer` on two half batches must end where one rank on the full batch ends -- for both ways of averaging the
gradients: the peer-memory all-reduce fused into the Adam kernel (gcnb_adam_tf_allreduce_f32) and the plain NCCL
all-reduce.  Dropout is off here: the counter-based masks are indexed by the LOCAL row, so a sharded batch draws
different masks than the full one by construction.
"""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dev):
    from gcn_fmri_decoding_b200 import synth
    from gcn_fmri_decoding_b200.models import cgcnn

    A, gs, perm, L = synth.brain_graph(4)
    return cgcnn(L=L, F=[32, 32], K=[5, 5], p=[4, 4], M=[512, 256, 22], channel=15, device=dev, seed=7,
                 regularization=5e-4, batch_size=32, perm=perm, n_input_vertices=360)


def _worker(rank, world, port, peer, out_path):
    import torch.distributed as dist

    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from gcn_fmri_decoding_b200 import synth
    from gcn_fmri_decoding_b200.train import FusedTrainer, shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 64
    x = torch.as_tensor(synth.bold_windows(B, seed=77), device=dev)
    y = torch.as_tensor(synth.labels(B, seed=77), device=dev)
    tr = FusedTrainer(_build(dev), distributed=True, use_cuda_graph=True, dropout=1.0, peer_allreduce=peer)
    lo, hi = shard(B, rank, world)
    for _ in range(3):
        tr.step(x[lo:hi].contiguous(), y[lo:hi].contiguous())
    torch.cuda.synchronize()
    gathered = [torch.empty_like(tr.flat_p) for _ in range(world)]
    dist.all_gather(gathered, tr.flat_p)
    if rank == 0:
        torch.save({"p": [g.cpu() for g in gathered], "kind": tr.allreduce_kind,
                    "err": getattr(tr, "_peer_error", None)}, out_path)
    dist.barrier()
    os._exit(0)  # the communicator lives in a captured graph: leave without tearing it down


@pytest.mark.parametrize("peer", [True, False])
def test_two_ranks_match_one_rank(tmp_path, peer):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp

    from gcn_fmri_decoding_b200 import synth
    from gcn_fmri_decoding_b200.train import FusedTrainer

    out_path = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), peer, out_path), nprocs=2, join=True)
    res = torch.load(out_path)
    if peer:
        assert res["kind"].startswith("peer-memory"), (res["kind"], res["err"])
    else:
        assert res["kind"] == "nccl"
    p0, p1 = res["p"]
    assert torch.equal(p0, p1)  # rank-ordered sums: the replicas are bit-identical
    dev = torch.device("cuda", 0)
    B = 64
    x = torch.as_tensor(synth.bold_windows(B, seed=77), device=dev)
    y = torch.as_tensor(synth.labels(B, seed=77), device=dev)
    ref = FusedTrainer(_build(dev), distributed=False, use_cuda_graph=True, dropout=1.0)
    start = ref.flat_p.clone()
    big = None
    for _ in range(3):
        ref.step(x, y)
        torch.cuda.synchronize()
        g = ref.flat_g.abs()
        sel = g > 1e-3 * float(g.max())
        big = sel if big is None else (big & sel)
    moved = (ref.flat_p - start).abs()
    assert int(big.sum()) > 1000 and float(moved[big].max()) > 1e-3       # the steps really moved the weights
    assert float((ref.flat_p.cpu() - p0)[big.cpu()].abs().max()) <= 2e-4  # three steps of size 1e-3
