"""Package host code (graphs/graclus/synth) against vectors produced by the reference's functions."""
import hashlib
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN
from gcn_fmri_decoding_b200 import graclus, graphs, synth
from oracle import ref_loader


def same_csr(a, b):
    a = sp.csr_matrix(a)
    b = sp.csr_matrix(b)
    a.sort_indices()
    b.sort_indices()
    return (a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.indptr, b.indptr)
            and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data))


def test_compute_perm_known_answer():
    """The reference's only executable check (coarsening.py:217-218)."""
    z = np.load(os.path.join(GOLDEN, "ref_fourier_perm.npz"))
    got = graclus.compute_perm([z["kat_parents0"], z["kat_parents1"]])
    assert got == [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]


@pytest.mark.parametrize("levels", [1, 4])
def test_pinned_graph_bit_exact(levels, graph_l1, graph_l4):
    ref = {1: graph_l1, 4: graph_l4}[levels]
    A, gs, perm, L = synth.brain_graph(levels)
    assert same_csr(A, ref["A"])
    assert np.array_equal(np.asarray(perm, np.int64), ref["perm"])
    assert [l.shape[0] for l in L] == ref["sizes"].tolist()
    for mine, theirs, theirs_t in zip(L, ref["L"], ref["Lt"]):
        assert same_csr(mine, theirs)
        assert same_csr(graphs.rescale_L(mine, 2), theirs_t)
    if levels == 4:
        assert [l.shape[0] for l in L] == [400, 200, 100, 50, 25]
        assert [graphs.rescale_L(l).nnz for l in L] == [3684, 1758, 898, 486, 266]
        assert hashlib.sha1(np.asarray(perm, np.int64).tobytes()).hexdigest()[:16] == "0d66436566ae4d31"
        assert str(ref["z"]["perm_sha1"]) == "0d66436566ae4d31"


def test_rescale_has_no_side_effects(graph_l4):
    L = graph_l4["L"][1].copy()
    before = L.data.copy()
    graphs.rescale_L(L, 2)
    assert np.array_equal(L.data, before)


def test_perm_data_3d_matches_reference_vector(graph_l4):
    z = np.load(os.path.join(GOLDEN, "ref_fourier_perm.npz"))
    out = graclus.perm_data_3d(z["x"], graph_l4["perm"])
    assert out.dtype == np.float64 and np.array_equal(out, z["x_perm"])
    fake = graph_l4["perm"] >= 360
    assert fake.sum() == 40 and np.all(out[:, fake, :] == 0)
    x = z["x"]
    assert graclus.perm_data_3d(x, None) is x
    with pytest.raises(ValueError):
        graclus.perm_data_3d(z["x"], np.arange(10))


def test_fourier_reconstructs(graph_l4):
    L = graph_l4["L"][2]
    lamb, U = graphs.fourier(L)
    assert np.allclose(U @ np.diag(lamb) @ U.T, L.toarray(), atol=1e-5)
    z = np.load(os.path.join(GOLDEN, "ref_fourier_perm.npz"))
    assert np.allclose(lamb, z["lamb"], atol=1e-5)


def test_csr_arrays_and_transpose(graph_l4):
    Lt = graph_l4["Lt"][0]
    rp, ci, v = graphs.csr_arrays(Lt)
    assert rp.dtype == np.int32 and ci.dtype == np.int32 and v.dtype == np.float32
    assert rp[-1] == 3684 and np.all(np.diff(rp) >= 0)
    for r in range(Lt.shape[0]):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)
    rpt, cit, vt = graphs.csr_arrays(Lt, transpose=True)
    T = sp.csr_matrix((vt, cit, rpt), shape=Lt.shape)
    assert (T != Lt.T).nnz == 0
    assert abs(Lt - Lt.T).max() > 0  # not bit-symmetric: the explicit adjoint matters


def test_synthetic_windows_and_sphere():
    x = synth.bold_windows(64)
    assert x.shape == (64, 360, 15) and x.dtype == np.float32
    assert abs(x.std() - 1) < 0.05
    y = synth.labels(1000)
    assert y.min() == 0 and y.max() == 20
    L = synth.fibonacci_sphere_graph(2000, 6)
    assert L.shape == (2000, 2000) and L.dtype == np.float32


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources only exist in the build container")
@pytest.mark.parametrize("levels", [0, 2, 6])
def test_live_reference_other_levels(levels):
    """Container-only: levels not in the fixtures, against the reference executed now."""
    import contextlib
    import io

    graph, coarsening = ref_loader.load()
    rng = np.random.RandomState(1234)
    z = rng.randn(360, 3).astype(np.float32)
    dist, idx = graph.distance_sklearn_metrics(z, k=8, metric="euclidean")
    A = graph.adjacency(dist, idx)
    np.random.seed(1234)
    A = graph.replace_random_edges(A, 0.01)
    with contextlib.redirect_stdout(io.StringIO()):
        gr, perm = coarsening.coarsen(A, levels=levels, self_connections=False)
    A2, gs2, perm2, L2 = synth.brain_graph(levels)
    assert same_csr(A, A2)
    assert (perm is None and perm2 is None) or list(perm) == list(perm2)
    for g, l2 in zip(gr, L2):
        assert same_csr(graph.laplacian(g, normalized=True), l2)
