"""The oracle against the reference's OWN layer code, executed without TensorFlow: ``cgcnn``'s methods are compiled from
``/root/reference/lib_new/models_gcn.py`` with ``tf`` bound to a NumPy stand-in for the ~25 ops they call
(``oracle/tf_shim.py``) and run on the same inputs and weights as ``oracle/layers_np.py``.  This pins the TRANSCRIPTION --
op order, every transpose / reshape / concat, the weight row order ``f*K + k``, the variable names and shapes the
reference creates, which variables enter the L2 term -- to the reference source; what stays unpinned is TensorFlow's
implementation of those ops.  Build container only (/root/reference does not exist on the GPU box)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import layers_np as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")

CASES = [  # filter, brelu, Laplacian levels available, F, K, p, head
    ("chebyshev5", "b1relu", [32, 32], [5, 5], [4, 4], [512, 256, 22]),          # BASELINE config 2
    ("chebyshev5", "b2relu", [32] * 6, [5] * 6, [1] * 6, [512, 256, 22]),        # config 1 (production network)
    ("chebyshev2", "b1relu", [32, 32], [2, 2], [4, 4], [512, 256, 22]),          # config 3a
    ("chebyshev5", "b1relu", [32, 32], [1, 1], [4, 4], [512, 256, 22]),          # config 3a, K = 1 ("firstorder")
    ("fourier", "b1relu", [32, 32], [0, 0], [4, 4], [512, 256, 22]),             # config 3b
    ("chebyshev5", "b2relu", [8, 8, 16], [20, 10, 10], [1, 4, 2], [16, 6]),      # mixed pooling, high order
]


def _variables(rng, Ls, filt, brelu, channel, F, K, M):
    v, fin = {}, channel
    for i, L in enumerate(Ls):
        Mi = L.shape[0]
        v["conv%d/weights" % (i + 1)] = (rng.randn(Mi, F[i], fin) if filt == "fourier" else rng.randn(fin * K[i], F[i])) * 0.2
        v["conv%d/bias" % (i + 1)] = rng.randn(1, Mi if brelu == "b2relu" else 1, F[i]) * 0.1 + 0.2
        fin = F[i]
    return v


@pytest.mark.parametrize("filt,brelu,F,K,p,Mfc", CASES)
def test_oracle_equals_the_reference_source_on_the_numpy_shim(graph_l1, graph_l4, filt, brelu, F, K, p, Mfc):
    g = graph_l1 if p == [1] * 6 else graph_l4
    L = g["L"]
    rng = np.random.RandomState(17)
    Ls = O.select_laplacians(L, p)
    channel, B = 15, 6
    var = _variables(rng, Ls, filt, brelu, channel, F, K, Mfc)
    width = -(-Ls[-1].shape[0] // p[-1])
    for i, m in enumerate(Mfc):
        scope = "logits" if i == len(Mfc) - 1 else "fc%d" % (i + 1)
        var[scope + "/weights"] = rng.randn(width, m) * 0.2
        var[scope + "/bias"] = rng.randn(m) * 0.1 + 0.2
        width = m
    var = {k: v.astype(np.float32) for k, v in var.items()}
    x = rng.randn(B, L[0].shape[0], channel).astype(np.float32)
    labels = rng.randint(0, Mfc[-1], B)

    # ---- the reference's code on the shim
    cgcnn, tf = ref_loader.load_cgcnn_on_shim(var)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = cgcnn("config", L, F, K, p, Mfc, filter=filt, brelu=brelu, pool="mpool1", channel=channel, regularization=5e-4,
                    dropout=0.5, batch_size=B)
        logits_ref = np.asarray(ref._inference(x.view(type(tf.constant(0.0))), 1))
        loss_ref, _ = ref.loss(logits_ref, labels, 5e-4)
    # the variables the reference created: names, order, shapes -- and all of them were supplied
    assert [r[0] for r in tf.requested] == list(var)
    assert all(tuple(var[name].shape) == shape for name, shape, _ in tf.requested)
    assert all(init == ("constant", 0.2) for name, _, init in tf.requested if name.endswith("bias"))
    n_reg = len(ref.regularizers)
    assert n_reg == (0 if filt == "chebyshev2" else len(p)) + 2 * len(Mfc)      # :583 (not regularised) vs :615 / :538 / :652-653

    # ---- the oracle
    params = [dict(W=var["conv%d/weights" % (i + 1)], b=var["conv%d/bias" % (i + 1)].reshape(
        (-1, F[i]) if brelu == "b2relu" else (F[i],)), K=K[i], p=p[i]) for i in range(len(p))]
    names = ["fc%d" % (i + 1) for i in range(len(Mfc) - 1)] + ["logits"]
    fcs = [(var[s + "/weights"], var[s + "/bias"]) for s in names]
    h = O.conv_stack(x, Ls, params, filter=filt, brelu=brelu, dtype=np.float32)
    logits = O.head(h, fcs, np.float32)
    assert logits.shape == logits_ref.shape == (B, Mfc[-1])
    assert np.array_equal(logits, logits_ref), float(np.abs(logits - logits_ref).max())   # same ops, same order: bit for bit
    regs = [pr["W"] for pr in params if filt != "chebyshev2"] + [v for Wb in fcs for v in Wb]
    assert abs(O.loss(logits.astype(np.float64), labels, regs, 5e-4) - float(loss_ref)) <= 1e-12 * abs(float(loss_ref))


@pytest.mark.parametrize("M,p", [(10, 4), (7, 2), (400, 4), (9, 8)])
def test_single_layer_pieces_on_the_shim(graph_l4, M, p):
    """b1relu / b2relu / mpool1 (ragged vertex counts included) called one by one on the reference source."""
    rng = np.random.RandomState(M + p)
    x = rng.randn(3, M, 5).astype(np.float32)
    for brelu, b in (("b1relu", rng.randn(1, 1, 5)), ("b2relu", rng.randn(1, M, 5))):
        cgcnn, tf = ref_loader.load_cgcnn_on_shim({"bias": b.astype(np.float32)})
        ref = cgcnn.__new__(cgcnn)
        ref.regularizers = []
        a_ref = np.asarray(getattr(ref, brelu)(tf.constant(x)))
        a = getattr(O, brelu)(x, b.astype(np.float32).reshape((M, 5) if brelu == "b2relu" else (5,)))
        assert np.array_equal(a, a_ref)
        assert np.array_equal(O.mpool1(a, p), np.asarray(ref.mpool1(tf.constant(a), p)))


def test_committed_golden_layer_vectors_are_the_reference_sources_output(graph_l4, layer_cases):
    """tests/golden/layer_cases.npz -- the vectors the GPU parity tests compare the kernels with -- against the reference's
    layer code on the shim: the fp32 as-run output ``y32`` bit for bit (Chebyshev cases; the spectral cases to fp32
    rounding, their eigenbasis comes from the host LAPACK of the generating run), the fp64 truth ``y64`` to 1e-12 with
    the shim in double precision."""
    for name, c in layer_cases.items():
        lvl, B, Fin, Fout, K, p = (int(v) for v in c["meta"])
        filt, brelu = str(c["kind"]).split("/")
        L = graph_l4["L"][lvl]
        M = L.shape[0]
        for dtype, key in ((np.float32, "y32"), (np.float64, "y64")):
            var = {"weights": np.asarray(c["W"], dtype), "bias": np.asarray(c["b"], dtype).reshape(1, -1, Fout)}
            cgcnn, tf = ref_loader.load_cgcnn_on_shim(var, dtype)
            ref = cgcnn.__new__(cgcnn)
            ref.regularizers, ref.initial = [], "normal"
            Lin = L if dtype == np.float32 else L.astype(np.float64)
            with contextlib.redirect_stdout(io.StringIO()):
                z = getattr(ref, filt)(tf.constant(c["x"]), Lin, Fout, K)
                y = np.asarray(ref.mpool1(getattr(ref, brelu)(z), p))
            assert y.shape == c[key].shape
            if filt == "fourier":
                assert np.abs(y - c[key]).max() <= 2e-4 * np.abs(c[key]).max(), (name, key)   # basis differs run to run
            elif dtype == np.float32:
                assert np.array_equal(y, c[key]), (name, float(np.abs(y - c[key]).max()))
            else:
                assert np.abs(y - c[key]).max() <= 1e-12 * np.abs(c[key]).max(), name


@pytest.mark.parametrize("filt,brelu,F,K,p,Mfc", CASES)
@pytest.mark.parametrize("keep", [1.0, 0.5])
def test_oracle_gradients_equal_autograd_of_the_reference_source(graph_l1, graph_l4, filt, brelu, F, K, p, Mfc, keep):
    """The training step: the reference's graph-building code (inference + loss, models_gcn.py:253-262, :658-682) run on
    the torch-backed stand-in and differentiated by torch.autograd in place of tf.gradients (:298), against the oracle's
    hand-written backward (``network_step``, fp64): loss to 1e-12, every gradient to 1e-10, with and without dropout
    masks.  chebyshev2: TensorFlow's py_func has no gradient, so the reference trains conv1 of a two-layer chebyshev2
    network with NO gradient at all -- asserted here; the oracle (and this repository) compute the
    full gradient, and what the reference does compute must agree."""
    g = graph_l1 if p == [1] * 6 else graph_l4
    L = [l.astype(np.float64) for l in g["L"]]       # both sides in double precision (graph.fourier's eigh included)
    rng = np.random.RandomState(23)
    Ls = O.select_laplacians(L, p)
    channel, B = 15, 4
    if len(p) == 6:
        F, Mfc = [6] * 6, [10, 8, 22]                      # narrower: fp64 autograd through six layers stays quick
    var = _variables(rng, Ls, filt, brelu, channel, F, K, Mfc)
    width = -(-Ls[-1].shape[0] // p[-1])
    for i, m in enumerate(Mfc):
        scope = "logits" if i == len(Mfc) - 1 else "fc%d" % (i + 1)
        var[scope + "/weights"] = rng.randn(width, m) * 0.2
        var[scope + "/bias"] = rng.randn(m) * 0.1 + 0.2
        width = m
    x = rng.randn(B, L[0].shape[0], channel)
    labels = rng.randint(0, Mfc[-1], B)
    masks = None if keep == 1.0 else [(rng.rand(B, m) < keep).astype(np.float64) for m in Mfc[:-1]]

    cgcnn, tf = ref_loader.load_cgcnn_on_shim(var, torch_autograd=True, dropout_masks=masks)
    torch = tf.torch
    with contextlib.redirect_stdout(io.StringIO()):
        ref = cgcnn("config", L, F, K, p, Mfc, filter=filt, brelu=brelu, pool="mpool1", channel=channel,
                    regularization=5e-4, dropout=keep, batch_size=B)
        logits = ref._inference(torch.tensor(x), keep)
        loss, _ = ref.loss(logits, labels, 5e-4)
    names = list(var)
    grads = dict(zip(names, torch.autograd.grad(loss, [tf.leaves[n] for n in names], allow_unused=True)))

    params = [dict(W=var["conv%d/weights" % (i + 1)], b=var["conv%d/bias" % (i + 1)].reshape(
        (-1, F[i]) if brelu == "b2relu" else (F[i],)), K=K[i], p=p[i]) for i in range(len(p))]
    fc_names = ["fc%d" % (i + 1) for i in range(len(Mfc) - 1)] + ["logits"]
    fcs = [(var[s + "/weights"], var[s + "/bias"]) for s in fc_names]
    val, cg, fg = O.network_step(x, labels, Ls, params, fcs, 5e-4, filter=filt, brelu=brelu, dtype=np.float64,
                                 dropout_masks=masks, keep=keep)
    assert abs(val - float(loss)) <= 1e-12 * abs(float(loss))

    def close(a, b):
        b = b.numpy()
        return np.abs(np.asarray(a).reshape(b.shape) - b).max() <= 1e-10 * max(np.abs(b).max(), 1e-30)

    for i, s in enumerate(fc_names):
        assert close(fg[i][0], grads[s + "/weights"]) and close(fg[i][1], grads[s + "/bias"]), s
    last = len(p) - 1
    for i in range(len(p)):
        gw, gb = grads["conv%d/weights" % (i + 1)], grads["conv%d/bias" % (i + 1)]
        if filt == "chebyshev2" and i < last:
            assert gw is None and gb is None            # the reference's back-propagation stops at the py_func of layer i+1
            continue
        assert close(cg[i]["dW"], gw) and close(cg[i]["db"], gb), i
