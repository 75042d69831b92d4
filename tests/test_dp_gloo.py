"""Data-parallel host logic on CPU: world_size 2 over gloo.

The conv kernels need a GPU, so a small stand-in network with the same interface
(``parameters()``, ``__call__(x, dropout=)``, ``loss()``) exercises exactly the code that is
device independent: batch sharding, the flat gradient buffer, the single all-reduce with
1/G scaling, the TF-style Adam update and weight broadcast at start.  Two ranks on half
batches must end up with the weights one process gets on the full batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcn_fmri_decoding_b200.train import TFAdam, Trainer, shard


class TinyNet(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w1 = torch.nn.Parameter(torch.randn(6, 5, generator=g) * 0.3)
        self.b1 = torch.nn.Parameter(torch.full((5,), 0.2))
        self.w2 = torch.nn.Parameter(torch.randn(5, 3, generator=g) * 0.3)
        self.regularization = 5e-4

    def forward(self, x, dropout=1.0):
        return torch.relu(x.mean(-1) @ self.w1 + self.b1) @ self.w2

    def loss(self, logits, labels):
        return torch.nn.functional.cross_entropy(logits, labels) + self.regularization * 0.5 * (self.w1 ** 2).sum()


def _data():
    g = torch.Generator().manual_seed(123)
    return torch.randn(8, 6, 4, generator=g), torch.randint(0, 3, (8,), generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, y = _data()
        model = TinyNet(seed=10 + rank)  # different init per rank: the trainer must broadcast rank 0's
        tr = Trainer(model)
        assert tr.world == world
        lo, hi = shard(x.shape[0], rank, world)
        for _ in range(3):
            tr.step(x[lo:hi], y[lo:hi])
        out[rank] = [p.detach().clone() for p in model.parameters()]
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_partitions_the_batch():
    for n, g in ((4096, 8), (4096, 4), (10, 3), (7, 8)):
        spans = [shard(n, r, g) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    assert shard(4096, 3, 8) == (1536, 2048)


def test_tf_adam_matches_closed_form():
    """TF-1.x Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps) (epsilon-hat form)."""
    p = torch.nn.Parameter(torch.tensor([1.0, -2.0, 0.5]))
    opt = TFAdam([p], lr=1e-3)
    ref = p.detach().clone().double().numpy()
    m = np.zeros(3)
    v = np.zeros(3)
    rng = np.random.RandomState(0)
    for t in range(1, 6):
        g = rng.randn(3)
        opt.step([torch.tensor(g, dtype=torch.float32)])
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        ref -= 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        assert np.allclose(p.detach().numpy(), ref, rtol=1e-5, atol=1e-7)


@pytest.mark.timeout(120)
def test_two_ranks_equal_one_process_on_the_full_batch():
    x, y = _data()
    single = TinyNet(seed=10)
    tr = Trainer(single, distributed=False)
    for _ in range(3):
        tr.step(x, y)
    want = [p.detach().clone() for p in single.parameters()]

    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(100)
            assert p.exitcode == 0
        got0, got1 = out[0], out[1]
    for a, b, w in zip(got0, got1, want):
        assert torch.equal(a, b)  # replicas stay identical
        assert torch.allclose(a, w, rtol=1e-5, atol=1e-6)  # mean of shard-means == full-batch mean


def test_momentum_zero_selects_decayed_sgd_like_the_reference():
    """models_gcn.py:283-292: momentum == 0 -> GradientDescentOptimizer on the staircase-decayed learning rate;
    anything else -> Adam(0.001).  Also: the state a CUDA-graph warm-up must restore is complete for both."""
    from gcn_fmri_decoding_b200.train import FusedTrainer, TFDecayedSGD

    x, y = _data()
    net = TinyNet(seed=3)
    net.momentum, net.learning_rate, net.decay_rate, net.decay_steps = 0, 0.1, 0.5, 2
    tr = Trainer(net, distributed=False)
    assert isinstance(tr.opt, TFDecayedSGD)
    ref = [p.detach().clone().double() for p in net.parameters()]
    twin = TinyNet(seed=3)
    for step in range(5):
        loss = twin.loss(twin(x), y)
        grads = torch.autograd.grad(loss, list(twin.parameters()))
        lr = 0.1 * 0.5 ** (step // 2)                         # staircase: global_step counts applied updates
        with torch.no_grad():
            for p, g in zip(twin.parameters(), grads):
                p -= lr * g
        tr.step(x, y)
        for a, b in zip(net.parameters(), twin.parameters()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), step
    assert float(tr.opt.global_step) == 5
    snap = tr._snapshot()
    tr.step(x, y)
    tr._restore(snap)
    assert float(tr.opt.global_step) == 5
    for a, b in zip(net.parameters(), snap[0]):
        assert torch.equal(a, b)
    # Adam: snapshot / restore puts back parameters, both moments and both power accumulators
    adam = Trainer(TinyNet(seed=4), distributed=False)
    assert isinstance(adam.opt, TFAdam) and len(adam.opt.state()) == 2 * 3 + 2
    adam.step(x, y)
    snap = adam._snapshot()
    before = [t.clone() for t in adam.opt.state()]
    adam.step(x, y)
    assert not torch.equal(adam.opt.b1_pow, before[-2])
    adam._restore(snap)
    assert all(torch.equal(a, b) for a, b in zip(adam.opt.state(), before))
    # no decay: plain SGD; decay without decay_steps is refused; the fused trainer refuses momentum == 0 loudly
    net2 = TinyNet(seed=5)
    net2.momentum, net2.learning_rate, net2.decay_rate, net2.decay_steps = 0, 0.05, 1, None
    p0 = [p.detach().clone() for p in net2.parameters()]
    g = torch.autograd.grad(net2.loss(net2(x), y), list(net2.parameters()))
    Trainer(net2, distributed=False).step(x, y)
    for a, b, gg in zip(net2.parameters(), p0, g):
        assert torch.allclose(a, b - 0.05 * gg, rtol=1e-6, atol=1e-8)
    with pytest.raises(ValueError):
        TFDecayedSGD(list(net2.parameters()), 0.1, None, 0.9)
    with pytest.raises(NotImplementedError, match="momentum"):
        FusedTrainer(net2)
