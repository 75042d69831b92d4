"""Model check of the handshake of ``k_adam_tf_allreduce`` (csrc/adam_allreduce.cu) -- the DESIGN, not the
CUDA code (that is tested on 2-8 GPUs by tests/test_dp_nccl.py and the scaling runs).

Per step every rank (1) writes its gradient buffer of the step's parity (backward kernels: many plain stores, modelled as
begin / end so that a reader can catch a half-written buffer), (2) raises one flag per peer with the exchange counter,
(3) waits until every peer's flag has reached the counter, (4) reads every rank's buffer of that parity.  There is NO
second handshake: the claim is that with TWO alternating buffers nobody ever reads a buffer that is being, or has
been, rewritten.  Seeded random schedulers interleave the ranks' atomic actions; the same model with ONE buffer must
fail (the check can see the hazard it is meant to exclude)."""
import random

import pytest


def run_schedule(world, steps, nbuf, rng):
    """One random interleaving; returns None or a description of the first violation."""
    G = [[("done", -1)] * nbuf for _ in range(world)]          # G[r][parity] = (state, step that wrote it)
    flags = [[0] * world for _ in range(world)]                 # flags[r][src] = exchange counter raised by src
    step = [0] * world
    pc = [0] * world                                            # program counter inside the step
    # actions of a step: 0 write-begin, 1 write-end, 2.. raise flag to each rank, then wait, then one read per rank
    n_actions = 2 + world + 1 + world
    while True:
        enabled = []
        for r in range(world):
            if step[r] >= steps:
                continue
            a = pc[r]
            if a == 2 + world and not all(flags[r][q] >= step[r] + 1 for q in range(world)):
                continue                                        # spinning on the peers' flags
            enabled.append(r)
        if not enabled:
            return None if all(s >= steps for s in step) else "deadlock at steps %r" % (step,)
        r = rng.choice(enabled)
        s, a, par = step[r], pc[r], step[r] % nbuf
        if a == 0:
            G[r][par] = ("writing", s)
        elif a == 1:
            G[r][par] = ("done", s)
        elif a < 2 + world:
            flags[a - 2][r] = s + 1                             # release store into rank (a-2)'s flag array
        elif a == 2 + world:
            pass                                                # all flags seen (acquire)
        else:
            q = a - (3 + world)
            if G[q][par] != ("done", s):
                return "rank %d, step %d read rank %d's buffer %d in state %r" % (r, s, q, par, G[q][par])
        pc[r] += 1
        if pc[r] == n_actions:
            pc[r], step[r] = 0, s + 1


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_two_alternating_buffers_need_no_second_handshake(world):
    rng = random.Random(1000 + world)
    for _ in range(600 if world < 8 else 150):
        assert run_schedule(world, steps=6, nbuf=2, rng=rng) is None


def test_the_model_sees_the_hazard_with_a_single_buffer():
    rng = random.Random(7)
    failures = [run_schedule(2, steps=6, nbuf=1, rng=rng) for _ in range(300)]
    assert any(f is not None and "read rank" in f for f in failures)
