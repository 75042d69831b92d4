"""Seeded synthetic inputs of the benchmark/test configurations.

The reference builds its brain graph from HCP data that is not available
(``/root/reference/model.py:62-147``); the pinned substitute is a 360-vertex
kNN graph of random 3-D points pushed through the same steps the reference
applies (``model.py:136-139``: ``replace_random_edges`` -> ``coarsen`` ->
``laplacian`` per level).  Windows are AR(1) "BOLD-like" series with unit
variance, which is what ``NDStandardScaler`` leaves behind (``utils.py:709-712``).
"""
from __future__ import annotations

import numpy as np

from . import graclus, graphs

N_ROI = 360
BLOCK_DURA = 15
N_CLASSES = 21  # labels in [0, 21); the logits have N_CLASSES + 1 = 22 columns


def brain_graph(levels, n_roi=N_ROI, k=8, noise=0.01, seed=1234):
    """Pinned synthetic brain graph: returns ``(A, graphs, perm, L)``.

    ``levels=1`` gives M=[372,186]; ``levels=4`` gives M=[400,200,100,50,25]
    with nnz(L~)=[3684,1758,898,486,266] (fixture ``tests/golden/graph_l4.npz``).
    """
    rng = np.random.RandomState(seed)
    z = rng.randn(n_roi, 3).astype(np.float32)
    dist, idx = graphs.distance_sklearn_metrics(z, k=k, metric="euclidean")
    A = graphs.adjacency(dist, idx)
    np.random.seed(seed)
    A = graphs.replace_random_edges(A, noise)
    gs, perm = graclus.coarsen(A, levels=levels, self_connections=False)
    L = [graphs.laplacian(g, normalized=True) for g in gs]
    return A, gs, perm, L


def fibonacci_sphere_graph(n=32492, k=6):
    """Vertex-level stand-in for a cortical mesh: kNN graph of a Fibonacci sphere.

    Returns the normalised Laplacian (CSR, fp32), nnz ~ 2e5 for the defaults
    (BASELINE config 5).  Neighbours come from a KD-tree; weights from
    ``graphs.adjacency``.
    """
    from scipy.spatial import cKDTree

    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    pts = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], 1).astype(np.float32)
    dist, idx = cKDTree(pts).query(pts, k=k + 1)
    A = graphs.adjacency(dist[:, 1:].astype(np.float32), idx[:, 1:])
    return graphs.laplacian(A.astype(np.float32), normalized=True)


def bold_windows(n, n_roi=N_ROI, t=BLOCK_DURA, seed=2024, rho=0.9):
    """``[n, n_roi, t]`` fp32 AR(1) windows, unit variance per entry."""
    rng = np.random.RandomState(seed)
    x = np.empty((n, n_roi, t), np.float32)
    x[:, :, 0] = rng.randn(n, n_roi)
    s = np.float32(np.sqrt(1 - rho * rho))
    for i in range(1, t):
        x[:, :, i] = np.float32(rho) * x[:, :, i - 1] + s * rng.randn(n, n_roi).astype(np.float32)
    return x


def labels(n, seed=2024):
    """Uniform task labels in ``[0, 21)`` (int64)."""
    return np.random.RandomState(seed + 1).randint(0, N_CLASSES, n).astype(np.int64)


def truncated_normal(rng, shape, std):
    """TF ``truncated_normal_initializer``: resample beyond two standard deviations (``models_gcn.py:333``)."""
    out = rng.randn(*shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.randn(int(bad.sum()))
        bad = np.abs(out) > 2
    return (out * std).astype(np.float32)
