"""Host-side graph operators that produce the operator inputs of the hot path.

These run once per experiment on the CPU (NumPy/SciPy) and define the *input
contract* of the CUDA layers: the normalised Laplacian ``L`` handed to
``cgcnn`` and the rescaled operator ``L~ = L/(lmax/2) - I`` the kernels
consume.  ``adjacency``, ``replace_random_edges``, ``laplacian``, ``lmax`` and
``rescale_L`` are TRANSLITERATED FROM the reference's ``lib_new/graph.py``
(``/root/reference/lib_new/graph.py:9-76`` kNN helpers, ``:79-98`` laplacian,
``:110-128`` fourier, ``:146-152`` rescale_L) statement by statement, on
purpose: they must consume the global NumPy RNG and round in exactly the
reference's order to reproduce the pinned graph bit for bit.  They are short
input-contract functions, not part of the work claimed by this repository, and
are checked bit-for-bit against the reference in ``tests/test_host_graph.py``
through the committed golden fixture (the reference itself cannot travel to the
GPU box).

Nothing here is on the timed path; no arithmetic of the layers themselves is
done on the host.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------- kNN graph
def distance_sklearn_metrics(z, k=4, metric="euclidean"):
    """k nearest neighbours of every row of ``z`` (excluding itself).

    Same contract as reference ``graph.py:9-17``: returns ``(dist[M,k], idx[M,k])``
    sorted by increasing distance.  The pairwise distances come from the same
    scikit-learn routine so that the pinned synthetic graph is reproducible to
    the last bit.
    """
    import sklearn.metrics

    d = sklearn.metrics.pairwise.pairwise_distances(z, metric=metric, n_jobs=-2)
    order = np.argsort(d)[:, 1 : k + 1]
    d.sort()
    return d[:, 1 : k + 1], order


def adjacency(dist, idx):
    """Symmetric Gaussian-weighted kNN adjacency (reference ``graph.py:19-45``).

    ``w_ij = exp(-d_ij^2 / sigma^2)`` with ``sigma`` the mean distance to the
    k-th neighbour; the graph is made undirected by keeping the larger of
    ``w_ij`` and ``w_ji``.
    """
    dist = np.asarray(dist)
    M, k = dist.shape
    if idx.shape != (M, k):
        raise ValueError("dist and idx must both be [M, k]")
    if dist.min() < 0:
        raise ValueError("distances must be non-negative")
    sigma2 = np.mean(dist[:, -1]) ** 2
    w = np.exp(-(dist ** 2) / sigma2)
    rows = np.arange(0, M).repeat(k)
    W = sp.coo_matrix((w.reshape(M * k), (rows, idx.reshape(M * k))), shape=(M, M))
    W.setdiag(0)
    # undirected: take the elementwise maximum of W and W^T
    take_t = W.T > W
    W = W - W.multiply(take_t) + W.T.multiply(take_t)
    assert W.nnz % 2 == 0
    assert abs(W - W.T).mean() < 1e-10
    return sp.csr_matrix(W)


def replace_random_edges(A, noise_level, rng=None):
    """Rewire ``noise_level`` of the undirected edges at random.

    Mirrors reference ``graph.py:48-76`` including its order of draws from the
    global NumPy generator (``permutation``, two ``randint``, ``uniform``), so
    that ``np.random.seed(s)`` before the call pins the result the same way.
    """
    rng = np.random if rng is None else rng
    M = A.shape[0]
    n = int(noise_level * A.nnz // 2)
    victims = rng.permutation(A.nnz // 2)[:n]
    new_r = rng.randint(0, M, n)
    new_c = rng.randint(0, M, n)
    _ = rng.uniform(0, 1, n)  # drawn (and unused) by the reference; keeps the stream aligned
    upper = sp.triu(A, format="coo")
    if upper.nnz < n:
        raise ValueError("not enough edges to rewire")
    A = A.tolil()
    for e, r, c in zip(victims, new_r, new_c):
        i, j = upper.row[e], upper.col[e]
        A[i, j] = 0
        A[j, i] = 0
        A[r, c] = 1
        A[c, r] = 1
    A.setdiag(0)
    A = A.tocsr()
    A.eliminate_zeros()
    return A


# --------------------------------------------------------------------------- Laplacians
def laplacian(W, normalized=True):
    """Graph Laplacian of a weight matrix (reference ``graph.py:79-98``).

    ``normalized``: ``I - D^-1/2 W D^-1/2`` with ``d += spacing(0)`` guarding
    isolated (fake) vertices; otherwise ``D - W``.  Result is CSR in W's dtype.
    """
    d = W.sum(axis=0)
    if not normalized:
        L = sp.diags(np.asarray(d).squeeze(), 0) - W
    else:
        d = d + np.spacing(np.array(0, W.dtype))
        d = 1 / np.sqrt(d)
        D = sp.diags(np.asarray(d).squeeze(), 0)
        I = sp.identity(d.size, dtype=W.dtype)
        L = I - D * W * D
    return sp.csr_matrix(L)


def lmax(L, normalized=True):
    """Upper bound of the spectrum (2 for a normalised Laplacian), ``graph.py:101-107``."""
    if normalized:
        return 2
    import scipy.sparse.linalg

    return scipy.sparse.linalg.eigsh(L, k=1, which="LM", return_eigenvectors=False)[0]


def rescale_L(L, lmax=2):
    """``L~ = L / (lmax/2) - I`` (reference ``graph.py:146-152``), without side effects.

    The reference divides and subtracts in place on a (shallow) copy
    (SURVEY D9); this version never touches its argument.  Explicit zeros the
    subtraction leaves on the diagonal are dropped, as SciPy does for the
    reference.
    """
    L = sp.csr_matrix(L, copy=True)
    M = L.shape[0]
    I = sp.identity(M, format="csr", dtype=L.dtype)
    L /= lmax / 2
    L = L - I
    L = sp.csr_matrix(L)
    return L


def fourier(L):
    """Graph Fourier basis ``(lambda, U)`` via dense ``eigh`` (reference ``graph.py:110-128``, algo 'eigh')."""
    lamb, U = np.linalg.eigh(L.toarray())
    return lamb, U


# --------------------------------------------------------------------------- CSR packing for the kernels
def csr_arrays(Lt, transpose=False):
    """(rowptr int32[M+1], col int32[nnz], val float32[nnz]) of a rescaled Laplacian.

    Rows are sorted by column (the order ``tf.sparse_reorder`` establishes at
    reference ``models_gcn.py:596``), values are cast to fp32 once here
    (SURVEY D8).  ``transpose=True`` returns the arrays of ``L~^T`` -- the
    explicit adjoint used by the backward recursion (``L~`` is not bit-symmetric,
    SURVEY A.4).
    """
    Lt = sp.csr_matrix(Lt.T if transpose else Lt, copy=True)
    Lt.sum_duplicates()
    Lt.sort_indices()
    return (
        np.ascontiguousarray(Lt.indptr, dtype=np.int32),
        np.ascontiguousarray(Lt.indices, dtype=np.int32),
        np.ascontiguousarray(Lt.data, dtype=np.float32),
    )
