"""A NumPy stand-in for the handful of TensorFlow-1.x ops the reference's layer code calls -- TEST INFRASTRUCTURE.

Purpose: run the reference's OWN source of ``cgcnn.chebyshev5 / chebyshev2 / fourier / filter_in_fourier / b1relu /
b2relu / mpool1 / fc / _inference`` and ``base_model.loss`` (``/root/reference/lib_new/models_gcn.py``) without
TensorFlow, so that the oracle (``oracle/layers_np.py``) is compared with the reference's code as written -- every
transpose, reshape, concat, the weight row order, the variable names and shapes, which variables join the L2 term --
instead of with a second transcription.  What this does NOT pin is TensorFlow's implementation of the ops below; each
is restated from its documented behaviour in a line or two of NumPy (the only non-trivial one is ``max_pool`` with SAME
padding).  Used by ``oracle/ref_loader.load_cgcnn_on_shim`` in the build container only (needs /root/reference).

Ops provided (call sites in models_gcn.py): ``SparseTensor`` / ``sparse_reorder`` / ``sparse_tensor_dense_matmul``
(:593-596, :605-608), ``transpose`` / ``reshape`` / ``expand_dims`` / ``concat`` / ``squeeze`` (:598-613, :633-637),
``matmul`` (:516-527, :616, :654), ``py_func`` (:578), ``constant`` (:536), ``nn.relu`` (:623, :655), ``nn.max_pool``
(:635), ``nn.dropout`` (:677), ``reduce_mean`` (:673, :259), ``nn.sparse_softmax_cross_entropy_with_logits`` (:258),
``nn.l2_loss`` (:345), ``add_n`` (:261), ``get_variable`` + ``variable_scope`` (:343, :351, :662), initialisers (:333-337),
and no-op ``name_scope`` / ``summary`` / ``train.ExponentialMovingAverage`` / ``control_dependencies`` / ``identity``.
"""
from __future__ import annotations

import contextlib
import types

import numpy as np
import scipy.sparse


class Tensor(np.ndarray):
    """ndarray with ``get_shape()`` (static shapes, as ``x.get_shape()`` at models_gcn.py:588)."""

    def get_shape(self):
        return tuple(int(s) for s in self.shape)


def _t(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(Tensor)


class _Sparse:
    def __init__(self, indices, values, dense_shape):
        indices = np.asarray(indices)
        self.csr = scipy.sparse.csr_matrix((np.asarray(values), (indices[:, 0], indices[:, 1])), shape=tuple(dense_shape))


class Shim:
    """One instance per graph build: holds the variables handed in (``get_variable`` looks them up by scoped name and
    checks the requested shape) and records what the reference code asked for."""

    float32, int32, int64 = np.float32, np.int32, np.int64

    def __init__(self, variables, dtype=np.float32):
        self.variables = {k: np.asarray(v, dtype) for k, v in variables.items()}
        self.dtype = dtype
        self.scopes = []
        self.requested = []            # (scoped name, shape, initialiser tag) in creation order

        # ---- tf.nn
        def max_pool(x, ksize, strides, padding):
            assert padding == "SAME" and list(ksize) == list(strides) and ksize[0] == ksize[2] == ksize[3] == 1
            p = int(ksize[1])
            N, M, F, one = x.shape
            Mo = -(-M // p)                       # SAME: ceil(M / stride) outputs
            pad = (Mo - 1) * p + p - M            # total padding, pad // 2 of it in front; padding never wins the max
            xp = np.full((N, Mo * p, F, one), -np.inf, x.dtype)
            xp[:, pad // 2: pad // 2 + M] = x
            return _t(xp.reshape(N, Mo, p, F, one).max(axis=2))

        def xent(logits=None, labels=None):
            z = np.asarray(logits, np.float64)
            z = z - z.max(axis=1, keepdims=True)
            logp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
            return _t(-logp[np.arange(len(labels)), np.asarray(labels)])

        def dropout(x, keep_prob):
            assert float(keep_prob) == 1.0, "the shim evaluates inference graphs (keep probability 1)"
            return x

        self.nn = types.SimpleNamespace(
            relu=lambda x: _t(np.maximum(x, 0)), max_pool=max_pool, dropout=dropout,
            sparse_softmax_cross_entropy_with_logits=xent,
            l2_loss=lambda v: np.sum(np.asarray(v, np.float64) ** 2) / 2)
        self.summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)

        class _EMA:
            def __init__(self, decay):
                pass

            def apply(self, xs):
                return None

            def average(self, x):
                return x

        self.train = types.SimpleNamespace(ExponentialMovingAverage=_EMA)
        self.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
            variance_scaling_initializer=lambda factor, mode, uniform: ("variance_scaling", factor, mode, uniform)))

    # ---- graph bookkeeping
    @contextlib.contextmanager
    def variable_scope(self, name):
        self.scopes.append(name)
        try:
            yield
        finally:
            self.scopes.pop()

    @contextlib.contextmanager
    def name_scope(self, name):
        yield

    @contextlib.contextmanager
    def control_dependencies(self, ops):
        yield

    def identity(self, x, name=None):
        return x

    def truncated_normal_initializer(self, mean, stddev):
        return ("truncated_normal", mean, stddev)

    def constant_initializer(self, value):
        return ("constant", value)

    def get_variable(self, name, shape, dtype, initializer=None):
        scoped = "/".join(self.scopes + [name])
        self.requested.append((scoped, tuple(int(s) for s in shape), initializer))
        if scoped not in self.variables:
            raise KeyError("the reference code asks for variable %r, which was not supplied" % scoped)
        v = self.variables[scoped]
        assert tuple(v.shape) == tuple(int(s) for s in shape), (scoped, v.shape, shape)
        out = _t(v)
        out.op = types.SimpleNamespace(name=scoped)      # `var.op.name` feeds tf.summary.histogram (:346, :354)
        return out

    # ---- tensors
    def constant(self, value, dtype=None):
        return _t(value, dtype or self.dtype)

    def to_int64(self, x):
        return np.asarray(x).astype(np.int64)

    def transpose(self, x, perm=None):
        return _t(np.transpose(x, perm))

    def reshape(self, x, shape):
        return _t(np.reshape(np.ascontiguousarray(x), [int(s) for s in shape]))

    def expand_dims(self, x, axis):
        return _t(np.expand_dims(x, axis))

    def concat(self, values, axis):
        return _t(np.concatenate(values, axis=axis))

    def squeeze(self, x, axis):
        return _t(np.squeeze(x, axis=tuple(axis)))

    def matmul(self, a, b):
        return _t(np.matmul(a, b))

    def reduce_mean(self, x, axis=None):
        return _t(np.mean(x, axis=axis, dtype=np.asarray(x).dtype))

    def add_n(self, xs):
        return sum(xs)

    def SparseTensor(self, indices, values, dense_shape):
        return _Sparse(indices, np.asarray(values, self.dtype), dense_shape)

    def sparse_reorder(self, sp):
        sp.csr.sort_indices()             # row-major canonical order
        return sp

    def sparse_tensor_dense_matmul(self, sp, x):
        return _t(sp.csr.dot(np.asarray(x)))

    def py_func(self, func, inp, Tout):
        return [_t(func(*[np.asarray(a) for a in inp]))]


class TorchShim(Shim):
    """The same stand-in on torch tensors (CPU, float64): the reference's graph-building code then builds a torch
    autograd graph, and ``torch.autograd.grad`` plays the part of ``tf.gradients`` (models_gcn.py:298) -- the reference
    source differentiated by an independent engine, which is what the oracle's hand-written backward is compared with.
    ``py_func`` (chebyshev2, :578) passes values but no gradient, exactly as in TensorFlow.  ``dropout_masks``: one 0/1
    array per ``tf.nn.dropout`` call, applied as TF does (kept values scaled by 1 / keep_prob)."""

    def __init__(self, variables, dropout_masks=None):
        import torch

        self.torch = torch
        if not hasattr(torch.Tensor, "get_shape"):
            torch.Tensor.get_shape = lambda t: tuple(int(s) for s in t.shape)   # static shapes (:588)
        super().__init__({}, np.float64)
        self.leaves = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in variables.items()}
        self.variables = {k: v.detach().numpy() for k, v in self.leaves.items()}
        masks = list(dropout_masks or [])

        def max_pool(x, ksize, strides, padding):
            assert padding == "SAME" and list(ksize) == list(strides) and ksize[0] == ksize[2] == ksize[3] == 1
            p = int(ksize[1])
            N, M, F, one = x.shape
            Mo = -(-M // p)
            pad = Mo * p - M
            xp = torch.nn.functional.pad(x.permute(0, 2, 3, 1), (pad // 2, pad - pad // 2), value=float("-inf"))
            return torch.nn.functional.max_pool1d(xp.reshape(N, F * one, Mo * p), p, p).reshape(N, F, one, Mo).permute(0, 3, 1, 2)

        def dropout(x, keep_prob):
            if not masks:
                assert float(keep_prob) == 1.0, "keep probability < 1 needs dropout_masks"
                return x
            return x * torch.tensor(np.asarray(masks.pop(0), np.float64)) / float(keep_prob)

        self.nn = types.SimpleNamespace(
            relu=torch.relu, max_pool=max_pool, dropout=dropout,
            sparse_softmax_cross_entropy_with_logits=lambda logits=None, labels=None: torch.nn.functional.cross_entropy(
                logits, torch.as_tensor(np.asarray(labels), dtype=torch.long), reduction="none"),
            l2_loss=lambda v: (v ** 2).sum() / 2)

    def get_variable(self, name, shape, dtype, initializer=None):
        super().get_variable(name, shape, dtype, initializer)      # bookkeeping and the shape check
        out = self.leaves["/".join(self.scopes + [name])]
        out.op = types.SimpleNamespace(name="/".join(self.scopes + [name]))
        return out

    def constant(self, value, dtype=None):
        return self.torch.tensor(np.ascontiguousarray(np.asarray(value, np.float64)))

    def to_int64(self, x):
        return np.asarray(x).astype(np.int64)

    def transpose(self, x, perm=None):
        return x.permute(*(perm if perm is not None else reversed(range(x.dim()))))

    def reshape(self, x, shape):
        return x.reshape([int(s) for s in shape])

    def expand_dims(self, x, axis):
        return x.unsqueeze(axis)

    def concat(self, values, axis):
        return self.torch.cat(list(values), dim=axis)

    def squeeze(self, x, axis):
        for a in sorted(axis, reverse=True):
            x = x.squeeze(a)
        return x

    def matmul(self, a, b):
        return self.torch.matmul(a, b)

    def reduce_mean(self, x, axis=None):
        return x.mean() if axis is None else x.mean(dim=axis)

    def SparseTensor(self, indices, values, dense_shape):
        indices = np.asarray(indices)
        return self.torch.sparse_coo_tensor(indices.T.copy(), np.asarray(values, np.float64), tuple(dense_shape)).coalesce()

    def sparse_reorder(self, sp):
        return sp

    def sparse_tensor_dense_matmul(self, sp, x):
        return self.torch.sparse.mm(sp, x)

    def py_func(self, func, inp, Tout):
        return [self.torch.tensor(np.asarray(func(*[a.detach().numpy() for a in inp])))]   # values only: no gradient
