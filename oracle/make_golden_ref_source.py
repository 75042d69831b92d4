"""Golden vectors produced by the REFERENCE'S OWN SOURCE (no TensorFlow): ``cgcnn._inference`` and ``loss`` of
``/root/reference/lib_new/models_gcn.py`` run on the NumPy stand-in (fp32 logits, as the reference runs) and on the
torch-backed stand-in (fp64 loss and ``torch.autograd`` gradients in place of ``tf.gradients``) -- see
``oracle/tf_shim.py``.  Written to ``tests/golden/ref_source_steps.npz`` so that the pin travels to the GPU box, where
/root/reference does not exist: ``tests/test_oracle.py`` checks the oracle against these vectors everywhere.
Build container only; test infrastructure."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import layers_np as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

# name: filter, brelu, graph fixture, F, K, p, head, keep
CASES = {
    "config2": ("chebyshev5", "b1relu", "ref_graph_l4.npz", [32, 32], [5, 5], [4, 4], [64, 32, 22], 0.5),
    "config1": ("chebyshev5", "b2relu", "ref_graph_l1.npz", [8] * 6, [5] * 6, [1] * 6, [32, 16, 22], 1.0),
    "config3a": ("chebyshev2", "b1relu", "ref_graph_l4.npz", [16, 16], [2, 2], [4, 4], [32, 16, 22], 0.5),
}
# (no spectral case: the eigenbasis of the padded graph's degenerate spectrum depends on the host's LAPACK, so such a vector
# would not travel; the spectral filter is compared with the reference source live, tests/test_reference_on_shim.py)


def graph(fixture):
    import scipy.sparse as sp

    z = np.load(os.path.join(ROOT, "tests", "golden", fixture))
    n = len(z["sizes"])
    return [sp.csr_matrix((z["L%d_data" % i], z["L%d_indices" % i], z["L%d_indptr" % i]),
                          shape=tuple(int(v) for v in z["L%d_shape" % i])) for i in range(n)]


def main():
    out = {}
    for ci, (name, (filt, brelu, fixture, F, K, p, Mfc, keep)) in enumerate(CASES.items()):
        L = graph(fixture)
        Ls = O.select_laplacians(L, p)
        rng = np.random.RandomState(100 + ci)
        channel, B = 15, 4
        var, fin = {}, channel
        for i, Li in enumerate(Ls):
            Mi = Li.shape[0]
            W = rng.randn(Mi, F[i], fin) if filt == "fourier" else rng.randn(fin * K[i], F[i])
            var["conv%d/weights" % (i + 1)] = (W * 0.2).astype(np.float32)
            var["conv%d/bias" % (i + 1)] = (rng.randn(1, Mi if brelu == "b2relu" else 1, F[i]) * 0.1 + 0.2).astype(np.float32)
            fin = F[i]
        width = -(-Ls[-1].shape[0] // p[-1])
        for i, m in enumerate(Mfc):
            scope = "logits" if i == len(Mfc) - 1 else "fc%d" % (i + 1)
            var[scope + "/weights"] = (rng.randn(width, m) * 0.2).astype(np.float32)
            var[scope + "/bias"] = (rng.randn(m) * 0.1 + 0.2).astype(np.float32)
            width = m
        x = rng.randn(B, L[0].shape[0], channel).astype(np.float32)
        labels = rng.randint(0, Mfc[-1], B)
        masks = None if keep == 1.0 else [(rng.rand(B, m) < keep).astype(np.float64) for m in Mfc[:-1]]
        with contextlib.redirect_stdout(io.StringIO()):
            # fp32, inference graph (keep probability 1), NumPy stand-in
            cg, tf = ref_loader.load_cgcnn_on_shim(var)
            ref = cg("config", L, F, K, p, Mfc, filter=filt, brelu=brelu, pool="mpool1", channel=channel,
                     regularization=5e-4, dropout=keep, batch_size=B)
            logits32 = np.asarray(ref._inference(tf.constant(x), 1))
            # fp64 training step, torch stand-in + autograd
            cg, tf = ref_loader.load_cgcnn_on_shim(var, torch_autograd=True, dropout_masks=masks)
            L64 = [l.astype(np.float64) for l in L]
            ref = cg("config", L64, F, K, p, Mfc, filter=filt, brelu=brelu, pool="mpool1", channel=channel,
                     regularization=5e-4, dropout=keep, batch_size=B)
            loss, _ = ref.loss(ref._inference(tf.torch.tensor(x.astype(np.float64)), keep), labels, 5e-4)
            grads = tf.torch.autograd.grad(loss, [tf.leaves[k] for k in var], allow_unused=True)
        out[name + ".meta"] = np.array([filt, brelu, fixture, repr(F), repr(K), repr(p), repr(Mfc), repr(keep)])
        out[name + ".x"], out[name + ".labels"], out[name + ".logits32"] = x, labels, logits32
        out[name + ".loss64"] = np.float64(float(loss))
        for k, v in var.items():
            out[name + ".var." + k.replace("/", "__")] = v
        for k, gr in zip(var, grads):
            if gr is not None:
                out[name + ".grad." + k.replace("/", "__")] = gr.numpy()
        for i, m in enumerate(masks or []):
            out[name + ".mask%d" % i] = m.astype(np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "ref_source_steps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
