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
