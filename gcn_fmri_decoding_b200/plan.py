"""Device-resident operator plans: what the kernels need of a Laplacian, uploaded once.

``GraphPlan`` is the replacement of the graph-build-time block of the reference's
``chebyshev5`` (``/root/reference/lib_new/models_gcn.py:590-596``: csr copy ->
``graph.rescale_L(L, lmax=2)`` -> COO -> ``tf.SparseTensor`` -> ``sparse_reorder``).
``SpectralPlan`` replaces ``models_gcn.py:535-536`` (``graph.fourier`` -> ``U.T`` constant).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib, graphs


class GraphPlan:
    """Rescaled Laplacian ``L~`` (CSR) and its explicit transpose on the device, fp32/int32.

    ``L`` is the *un-rescaled* normalised Laplacian exactly as ``graph.laplacian`` returns it
    (fp32 or fp64, any SciPy sparse format); values are cast to fp32 once here.
    The transpose is stored explicitly because ``L~`` is not bit-symmetric.
    """

    def __init__(self, L, device="cuda", lmax=2, rescale=True):
        if not sp.issparse(L):
            raise ValueError("L must be a scipy sparse matrix")
        if L.shape[0] != L.shape[1]:
            raise ValueError("L must be square")
        Lt = graphs.rescale_L(L, lmax=lmax) if rescale else sp.csr_matrix(L)
        self.M = int(Lt.shape[0])
        self.device = torch.device(device)
        rp, ci, v = graphs.csr_arrays(Lt)
        rpt, cit, vt = graphs.csr_arrays(Lt, transpose=True)
        self.nnz = int(v.shape[0])
        self.max_degree = int(np.diff(rp).max()) if self.M > 0 else 0
        t = lambda a: torch.from_numpy(a).to(self.device)
        self.rowptr, self.col, self.val = t(rp), t(ci), t(v)
        self.rowptr_t, self.col_t, self.val_t = t(rpt), t(cit), t(vt)

    def tensors(self):
        return self.rowptr, self.col, self.val, self.rowptr_t, self.col_t, self.val_t


def csr_struct(rowptr, col, val):
    """``gcnb_csr`` for three device tensors (kept alive by the caller for the duration of the call)."""
    return _lib.GcnbCsr(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), rowptr.numel() - 1, val.numel())


class SpectralPlan:
    """``Ut`` = transposed eigenvector matrix of ``L`` as an fp32 device tensor ``[M, M]``."""

    def __init__(self, L=None, device="cuda", Ut=None):
        if Ut is None:
            _, U = graphs.fourier(L)
            Ut = np.ascontiguousarray(U.T).astype(np.float32)
        Ut = np.ascontiguousarray(Ut, dtype=np.float32)
        if Ut.ndim != 2 or Ut.shape[0] != Ut.shape[1]:
            raise ValueError("Ut must be [M, M]")
        self.M = int(Ut.shape[0])
        self.Ut = torch.from_numpy(Ut).to(device)
