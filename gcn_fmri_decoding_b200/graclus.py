"""Graclus multilevel coarsening and the binary-tree vertex ordering.

Host-side, runs once per experiment.  It defines, bit for bit, the vertex
order the ``mpool1`` kernel pools over (siblings adjacent, fake vertices
padded in), so it has to agree exactly with the reference
(``/root/reference/lib_new/coarsening.py``: ``coarsen :5-31``, ``metis :34-116``,
``metis_one_level :120-166``, ``compute_perm :168-215``, ``perm_data_3d :244-265``,
``perm_adjacency :267-294``).  Agreement is pinned by the reference's own
known-answer vector for ``compute_perm`` (``coarsening.py:217-218``) and by the
golden fixture ``tests/golden/graph_l4.npz`` generated from the reference.

The data permutation also exists on the device (``ops.perm_gather``); the NumPy
``perm_data_3d`` here is the drop-in host form.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def coarsen(A, levels, self_connections=False, verbose=False):
    """Coarsen ``A`` ``levels`` times; returns ``(graphs, perm)``.

    ``graphs[i]`` is the adjacency at level ``i`` re-ordered (and padded with
    isolated fake vertices) so that vertices ``2j, 2j+1`` of level ``i`` merge
    into vertex ``j`` of level ``i+1``; ``perm`` is the level-0 ordering to be
    applied to the data (``perm_data_3d``).
    """
    graphs, parents = metis(A, levels)
    perms = compute_perm(parents)
    for i, G in enumerate(graphs):
        M = G.shape[0]
        if not self_connections:
            G = G.tocoo()
            G.setdiag(0)
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = G.tocsr()
        G.eliminate_zeros()
        graphs[i] = G
        if verbose:
            print("level %d: %d vertices (%d fake), %d edges" % (i, G.shape[0], G.shape[0] - M, G.nnz // 2))
    return graphs, (perms[0] if levels > 0 else None)


def metis(W, levels, rid=None):
    """Greedy Graclus pairing repeated ``levels`` times -> ``(graphs, parents)``.

    ``parents[i][v]`` is the cluster (vertex of level ``i+1``) vertex ``v`` of
    level ``i`` belongs to.  The first visiting order is a permutation drawn
    with ``np.random.seed(1234)`` exactly as the reference does
    (``coarsening.py:55-57``); later levels visit by increasing weighted degree.
    """
    N = W.shape[0]
    if rid is None:
        np.random.seed(1234)
        rid = np.random.permutation(range(N))
    degree = W.sum(axis=0) - W.diagonal()
    graphs, parents = [W], []
    for _ in range(levels):
        weights = np.array(degree).squeeze()
        # entries grouped by row; np.argsort's default kind is kept on purpose: the
        # within-row order decides ties in the greedy matching
        r, c, v = sp.find(W)
        order = np.argsort(r)
        rr, cc, vv = r[order], c[order], v[order]
        cluster = metis_one_level(rr, cc, vv, rid, weights)
        parents.append(cluster)
        Nnew = int(cluster.max()) + 1
        W = sp.csr_matrix((vv, (cluster[rr], cluster[cc])), shape=(Nnew, Nnew))
        W.eliminate_zeros()
        graphs.append(W)
        degree = W.sum(axis=0)
        rid = np.argsort(np.array(W.sum(axis=0)).squeeze())
    return graphs, parents


def metis_one_level(rr, cc, vv, rid, weights):
    """One greedy matching pass (``coarsening.py:120-166``).

    Vertices are visited in ``rid`` order; an unmatched vertex ``t`` is paired
    with the unmatched neighbour ``n`` maximising ``w_tn (1/d_t + 1/d_n)``
    (strictly greater wins, so the first best neighbour in row order is kept),
    or stays a singleton.  ``rr`` must be sorted.
    """
    nnz = rr.shape[0]
    N = int(rr[nnz - 1]) + 1
    # Row extents exactly as the reference's single scan derives them (coarsening.py:134-139):
    # slots are the *distinct* values of rr in order, and the scan bumps a slot's length
    # before it notices the row changed -- so slot 0 is one entry too long (it also sees the
    # first entry of the next row) and the last slot one entry too short.  The pairing, and
    # therefore the pinned permutation, depends on this; it is reproduced, not "fixed".
    starts = np.flatnonzero(np.r_[True, rr[1:] > rr[:-1]])
    lengths = np.diff(np.r_[starts, nnz])
    lengths[0] += 1
    lengths[-1] -= 1
    rowstart = np.zeros(N, np.int64)
    rowlength = np.zeros(N, np.int64)
    rowstart[: starts.size] = starts
    rowlength[: starts.size] = lengths

    matched = np.zeros(N, bool)
    cluster = np.zeros(N, np.int32)
    count = 0
    for t in np.asarray(rid)[:N].tolist():
        if matched[t]:
            continue
        matched[t] = True
        best, wmax = -1, 0.0
        s = rowstart[t]
        for j in range(s, s + rowlength[t]):
            n = cc[j]
            if matched[n]:
                gain = 0.0
            else:
                # same expression, same evaluation order and dtypes as the reference
                gain = vv[j] * (1.0 / weights[t] + 1.0 / weights[n])
            if gain > wmax:
                wmax, best = gain, n
        cluster[t] = count
        if best > -1:
            cluster[best] = count
            matched[best] = True
        count += 1
    return cluster


def compute_perm(parents):
    """Vertex orderings that turn the cluster hierarchy into a complete binary tree.

    Returns one index list per level, finest first.  Indices ``>= len(parent)``
    denote fake vertices added so every coarse vertex has exactly two children
    (``coarsening.py:168-215``).
    """
    if len(parents) == 0:
        return []
    orderings = [list(range(int(max(parents[-1])) + 1))]
    for parent in reversed(parents):
        parent = np.asarray(parent)
        next_fake = len(parent)
        # children of each coarse vertex, in increasing fine index
        by_cluster = {}
        for child, c in enumerate(parent.tolist()):
            by_cluster.setdefault(c, []).append(child)
        layer = []
        for c in orderings[-1]:
            kids = list(by_cluster.get(c, []))
            if len(kids) > 2:
                raise ValueError("a cluster has more than two children")
            while len(kids) < 2:
                kids.append(next_fake)
                next_fake += 1
            layer.extend(kids)
        orderings.append(layer)
    M_last = len(orderings[0])
    for i, layer in enumerate(orderings):
        assert sorted(layer) == list(range(M_last * 2 ** i))
    return orderings[::-1]


def perm_adjacency(A, indices):
    """Pad ``A`` with isolated vertices to ``len(indices)`` and re-order it (``coarsening.py:267-294``)."""
    if indices is None:
        return A
    M = A.shape[0]
    Mnew = len(indices)
    if Mnew < M:
        raise ValueError("ordering shorter than the graph")
    A = A.tocoo()
    rank = np.argsort(indices)  # rank[old] = new position
    return sp.coo_matrix((A.data, (np.asarray(rank)[A.row], np.asarray(rank)[A.col])), shape=(Mnew, Mnew))


def perm_data(x, indices):
    """Re-order / zero-pad the vertex axis of ``x [N, M]`` (``coarsening.py:221-241``)."""
    if indices is None:
        return x
    return perm_data_3d(np.asarray(x)[:, :, None], indices)[:, :, 0]


def perm_data_3d(x, indices):
    """Re-order / zero-pad the vertex axis of ``x [N, M, F]`` (``coarsening.py:244-265``).

    ``out[:, i, :] = x[:, indices[i], :]`` for real vertices, 0 for fake ones
    (``indices[i] >= M``).  Returns float64 like the reference (``np.empty``
    default dtype); the device op ``perm_gather`` keeps fp32.
    """
    if indices is None:
        return x
    x = np.asarray(x)
    N, M, F = x.shape
    indices = np.asarray(indices)
    if len(indices) < M:
        raise ValueError("ordering shorter than the data")
    out = np.zeros((N, len(indices), F))
    real = indices < M
    out[:, real, :] = x[:, indices[real], :]
    return out
