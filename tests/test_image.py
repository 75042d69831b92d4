"""Host-side operator image (gcnb_cheb_image_build, include/gcnb200.h): decoding it must give back the rescaled
Laplacian bit for bit -- every block's entry list is the union of its four rows' neighbour sets with the exact weights,
padding steps gather the zero row with zero weights.  No GPU: the builder is host code."""
import numpy as np
import pytest


def _decode(img, M, p_eff):
    hdr = img[:64].view(np.uint32)
    assert hdr[0] == 0x494E4347 and hdr[2] == img.size
    ng = int(hdr[3])
    grp = img[64:64 + ng * 8].view(np.uint32).reshape(ng, 2)
    blk_off = 64 + ((ng * 8 + 15) // 16) * 16
    blk = img[blk_off:blk_off + ng * 8].view(np.uint16).reshape(ng, 4)
    assert sorted(int(b) for b in blk.ravel() if b != 0xFFFF) == list(range(M // 4))
    assert (np.diff(grp[:, 1].astype(np.int64)) <= 0).all() and (grp[:, 1] % 2 == 0).all()  # sorted, even step counts
    pairs = []
    for g in range(ng):
        off, steps = int(grp[g, 0]), int(grp[g, 1])
        for t in range(steps // 2):
            pair = img[off + t * 144:off + (t + 1) * 144]
            pairs.append((g, pair[:128].view(np.float32).reshape(2, 4, 4), pair[128:144].view(np.uint16).reshape(4, 2)))
    zero_row = max(int(c.max()) for _, _, c in pairs) >> 3  # the zero row is the last row of a state buffer
    BQ, log2p = zero_row // p_eff, int(np.log2(p_eff))
    row2v = {(v & (p_eff - 1)) * BQ + (v >> log2p): v for v in range(M)}
    dense = np.zeros((M, M), np.float32)
    for g, w, codes in pairs:
        for q in range(4):
            beta = int(blk[g, q])
            for jj in range(2):
                code = int(codes[q, jj])                             # 16-bit row code = gather code >> 4
                row = code >> 3
                assert code & 7 == row & 7                           # SWIZZLE_128B phase of the gathered row
                if row == zero_row:
                    assert not w[jj, q].any()
                    continue
                assert beta != 0xFFFF
                col = row2v[row]
                for i in range(4):
                    assert dense[beta * 4 + i, col] == 0
                    dense[beta * 4 + i, col] = w[jj, q, i]
    return dense


@pytest.mark.parametrize("level,B,Fin,Fout,K,p,adjoint", [(0, 512, 15, 32, 5, 4, 0), (2, 512, 32, 32, 5, 4, 0),
                                                          (2, 512, 32, 32, 5, 4, 1), (1, 64, 16, 32, 3, 2, 0),
                                                          (2, 7, 12, 8, 2, 1, 0)])
def test_image_decodes_to_the_operator(graph_l4, level, B, Fin, Fout, K, p, adjoint):
    from gcn_fmri_decoding_b200 import _lib, graphs

    Lr = graphs.rescale_L(graph_l4["L"][level], lmax=2)
    rp, ci, v = graphs.csr_arrays(Lr, transpose=bool(adjoint))
    M, nnz = Lr.shape[0], len(v)
    lib = _lib.lib()
    args = (B, M, nnz, Fin, Fout, K, p, adjoint)
    n = lib.gcnb_cheb_image_bytes(rp.ctypes.data, ci.ctypes.data, *args)
    assert n > 0 and n % 16 == 0
    img = np.zeros(n, np.uint8)
    _lib.check(lib.gcnb_cheb_image_build(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, *args, img.ctypes.data, n), "build")
    ref = np.asarray((Lr.T if adjoint else Lr).todense(), np.float32)
    assert np.array_equal(_decode(img, M, 1 if adjoint else p), ref)
    # wrong buffer size -> error code, message set
    assert lib.gcnb_cheb_image_build(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, *args, img.ctypes.data, n - 16) != 0
    assert "image" in _lib.last_error()


def test_image_unsupported_shapes_return_zero(graph_l4):
    from gcn_fmri_decoding_b200 import _lib, graphs

    lib = _lib.lib()
    Lr = graphs.rescale_L(graph_l4["L"][3], lmax=2)  # M = 50: not a multiple of four
    rp, ci, v = graphs.csr_arrays(Lr)
    assert lib.gcnb_cheb_image_bytes(rp.ctypes.data, ci.ctypes.data, 8, 50, len(v), 32, 32, 5, 2, 0) == 0
    Lr = graphs.rescale_L(graph_l4["L"][0], lmax=2)
    rp, ci, v = graphs.csr_arrays(Lr)
    assert lib.gcnb_cheb_image_bytes(rp.ctypes.data, ci.ctypes.data, 8, 400, len(v), 64, 32, 5, 4, 0) == 0  # Fin > 32
    bad = rp.copy()
    bad[5] = bad[6] + 1  # row pointers must not decrease
    assert lib.gcnb_cheb_image_bytes(bad.ctypes.data, ci.ctypes.data, 8, 400, len(v), 15, 32, 5, 4, 0) == 0


def test_production_layer_shape_has_an_image_and_a_stack_kernel(graph_l1):
    """The 372-vertex, 32 -> 32, p = 1 layers of the production network (config 1) must fit an operator image beside their
    state (16-bit row codes; the 32-bit format did not) -- the layer-stack kernel needs it.  Host-side queries only."""
    import ctypes as C

    from gcn_fmri_decoding_b200 import _lib, graphs

    lib = _lib.lib()
    Lr = graphs.rescale_L(graph_l1["L"][0], lmax=2)
    rp, ci, v = graphs.csr_arrays(Lr)
    M, nnz = Lr.shape[0], len(v)
    assert M == 372
    n = lib.gcnb_cheb_image_bytes(rp.ctypes.data, ci.ctypes.data, 128, M, nnz, 32, 32, 5, 1, 0)
    assert 0 < n < 64 * 1024 and n % 16 == 0
    img = np.zeros(n, np.uint8)
    _lib.check(lib.gcnb_cheb_image_build(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, 128, M, nnz, 32, 32, 5, 1, 0,
                                         img.ctypes.data, n), "build")
    assert np.array_equal(_decode(img, M, 1), np.asarray(Lr.todense(), np.float32))
    # the support query looks at the shape and at the presence / size of the image only (no device access)
    csr = _lib.GcnbCsr(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, M, nnz, img.ctypes.data, n)
    assert lib.gcnb_cheb_stack_supported(C.byref(csr), 128, 32, 5, 5) == 1
    assert lib.gcnb_cheb_stack_supported(C.byref(csr), 128, 32, 5, 9) == 0      # at most eight layers
    assert lib.gcnb_cheb_stack_supported(C.byref(csr), 128, 16, 5, 5) == 0      # 32 -> 32 filters only
    bare = _lib.GcnbCsr(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, M, nnz, None, 0)
    assert lib.gcnb_cheb_stack_supported(C.byref(bare), 128, 32, 5, 5) == 0     # needs the image


def _tf32_rna(x):
    """cvt.rna.tf32.f32 on the host: round to nearest, ties away from zero, 10 mantissa bits kept."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _bf16_rne(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


@pytest.mark.parametrize("Fin,Fout,K", [(32, 32, 5), (15, 32, 5), (12, 8, 2), (32, 4, 1)])
def test_tap_image_is_the_split_of_the_weights(Fin, Fout, K):
    """gcnb_cheb_tap_image_build (host code, include/gcnb200.h): three planes in the kernel's shared-memory layout --
    tf32 hi = rna(W), tf32 lo = rna(W - hi), bf16 = rne(W) -- of W[fin*K + k, fout] (models_gcn.py:613-616), zero in
    the padding of the filter dimensions.  hi + lo carries W to 2^-21."""
    from gcn_fmri_decoding_b200 import _lib

    lib = _lib.lib()
    n = lib.gcnb_cheb_tap_image_bytes(Fin, Fout, K)
    FP = 16 if Fin <= 16 else 32
    plane = K * FP * 32 * 4
    assert n == 2 * plane + plane // 2
    W = (np.random.RandomState(Fin * 100 + Fout + K).randn(Fin * K, Fout) * 0.3).astype(np.float32)
    W[0, 0], W[-1, -1] = 0.0, np.float32(1 + 2.0 ** -11)  # an exact zero and a rounding tie of the tf32 split
    img = np.full(n, 0xAB, np.uint8)
    _lib.check(lib.gcnb_cheb_tap_image_build(W.ctypes.data, Fin, Fout, K, img.ctypes.data, n), "tap image")
    kk, o = np.meshgrid(np.arange(K * FP), np.arange(32), indexing="ij")
    off32 = ((kk >> 2) * 4 + (o >> 3)) * 128 + (o & 7) * 16 + (kk & 3) * 4
    off16 = ((kk >> 3) * 4 + (o >> 3)) * 128 + (o & 7) * 16 + (kk & 7) * 2
    assert len(np.unique(off32)) == off32.size and off32.max() == plane - 4       # the layout is a permutation
    assert len(np.unique(off16)) == off16.size and off16.max() == plane // 2 - 2
    hi = img[:plane].view(np.float32)[off32 // 4]
    lo = img[plane:2 * plane].view(np.float32)[off32 // 4]
    bf = img[2 * plane:].view(np.uint16)[off16 // 2]
    k, f = kk // FP, kk % FP
    want = np.where((f < Fin) & (o < Fout), W[np.minimum(f, Fin - 1) * K + k, np.minimum(o, Fout - 1)], np.float32(0))
    want = want.astype(np.float32)
    assert np.array_equal(hi.view(np.uint32), _tf32_rna(want).view(np.uint32))
    assert np.array_equal(lo.view(np.uint32), _tf32_rna(want - _tf32_rna(want)).view(np.uint32))
    assert np.array_equal(bf, _bf16_rne(want))
    assert np.abs((hi.astype(np.float64) + lo) - want).max() <= 2.0 ** -21 * np.abs(want).max()
    assert lib.gcnb_cheb_tap_image_bytes(8, 32, 5) == 0 and lib.gcnb_cheb_tap_image_bytes(32, 30, 5) == 0
    assert lib.gcnb_cheb_tap_image_build(W.ctypes.data, Fin, Fout, K, img.ctypes.data, n - 16) != 0  # wrong size: refused


@pytest.mark.parametrize("seed", range(24))
def test_image_of_random_operators_decodes_exactly(seed):
    """Seeded sweep of the host image builder over operators that are NOT Graclus Laplacians: random sparsity (empty
    rows, rows up to 40 entries, entries on the diagonal), sizes up to what shared memory can take, every pooling size
    and the adjoint flag.  Whenever the library offers an image it must decode to the operator exactly; when it
    declines (returns 0 bytes) that is a size decision, never an error."""
    import scipy.sparse as sp

    from gcn_fmri_decoding_b200 import _lib

    rng = np.random.RandomState(100 + seed)
    M = int(rng.choice([8, 36, 100, 252, 372, 400, 512, 1000]))
    p = int(rng.choice([q for q in (1, 2, 4, 8) if M % q == 0]))
    Fin = int(rng.choice([9, 15, 16, 17, 32]))
    Fout = int(rng.choice([4, 8, 32]))
    K = int(rng.randint(1, 7))
    adjoint = int(rng.randint(0, 2))
    deg = rng.randint(0, int(rng.choice([3, 10, 25])) + 1, M)
    deg[rng.rand(M) < 0.1] = 0                                         # empty rows (the reference's fake vertices)
    rows = np.repeat(np.arange(M), deg)
    cols = rng.randint(0, M, len(rows))
    A = sp.csr_matrix((rng.randn(len(rows)).astype(np.float32), (rows, cols)), shape=(M, M))
    A.sum_duplicates()
    A.sort_indices()
    A.data[A.data == 0] = 1.0
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    lib = _lib.lib()
    args = (8, M, len(v), Fin, Fout, K, p, adjoint)
    n = lib.gcnb_cheb_image_bytes(rp.ctypes.data, ci.ctypes.data, *args)
    if not adjoint and M <= 400 and len(v) <= 4000:
        assert n > 0, "small forward operators must get an image"  # (adjoint mode swaps Fin / Fout: narrow Fout declines)
    if n == 0:
        return
    assert n % 16 == 0
    img = np.zeros(n, np.uint8)
    _lib.check(lib.gcnb_cheb_image_build(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, *args, img.ctypes.data, n), "build")
    if len(v) == 0:
        return
    assert np.array_equal(_decode(img, M, 1 if adjoint else p), np.asarray(A.todense(), np.float32))


def test_three_term_split_error_model():
    """Numerics of the tcgen05 contraction, emulated on the CPU: z = trunc_tf32(X) tf32(W) +
    trunc_tf32(X) tf32(W - tf32(W)) + bf16(X - trunc_tf32(X)) bf16(W) with exact products and wide accumulation.  The
    terms left out are (X - X_hi)(W - bf16 W), the bf16 rounding of the remainder and W - W_hi - W_lo: at most
    (2^-10 2^-9 + 2^-10 2^-9 + 2^-22) |x||w| < 4.1e-6 |x||w| per product -- two orders below the 1e-4 parity tolerance,
    which is why one bf16 remainder term is enough.  Checked on operands shaped like conv 1 of config 2 (75 -> 32)."""
    rng = np.random.RandomState(0)
    X = (rng.randn(4096, 75) * np.exp(rng.randn(4096, 75))).astype(np.float32)      # wide dynamic range
    W = (rng.randn(75, 32) * 0.2).astype(np.float32)
    x_hi = (X.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)              # what kind::tf32 reads of X
    x_lo = (_bf16_rne(X - x_hi).astype(np.uint32) << 16).view(np.float32)
    w_hi = _tf32_rna(W)
    w_lo = _tf32_rna(W - w_hi)
    w_b = (_bf16_rne(W).astype(np.uint32) << 16).view(np.float32)
    f8 = np.float64
    z = x_hi.astype(f8) @ w_hi.astype(f8) + x_hi.astype(f8) @ w_lo.astype(f8) + x_lo.astype(f8) @ w_b.astype(f8)
    exact = X.astype(f8) @ W.astype(f8)
    bound = 4.1e-6 * (np.abs(X).astype(f8) @ np.abs(W).astype(f8))
    assert np.all(np.abs(z - exact) <= bound)
    assert np.abs(z - exact).max() / np.abs(exact).max() < 2e-6                        # the parity norm: far below 1e-4
    one_pass = x_hi.astype(f8) @ w_hi.astype(f8)                                        # plain TF32 would NOT do:
    assert np.abs(one_pass - exact).max() / np.abs(exact).max() > 1e-4


def test_host_builders_are_reentrant(graph_l4):
    """include/gcnb200.h promises re-entrancy (no global mutable state besides the thread-local error string): eight
    threads build operator and tap images concurrently (ctypes releases the GIL) and get the single-threaded bytes;
    an error in one thread does not leak into another thread's error string."""
    from concurrent.futures import ThreadPoolExecutor

    from gcn_fmri_decoding_b200 import _lib, graphs

    lib = _lib.lib()
    jobs = []
    for level, (Fin, Fout, K, p) in enumerate([(15, 32, 5, 4), (32, 32, 5, 2), (32, 32, 3, 4)]):
        Lr = graphs.rescale_L(graph_l4["L"][level], lmax=2)
        rp, ci, v = graphs.csr_arrays(Lr)
        jobs.append((rp, ci, v, (64, Lr.shape[0], len(v), Fin, Fout, K, p, 0)))

    def build(job):
        rp, ci, v, args = job
        n = lib.gcnb_cheb_image_bytes(rp.ctypes.data, ci.ctypes.data, *args)
        img = np.zeros(n, np.uint8)
        assert lib.gcnb_cheb_image_build(rp.ctypes.data, ci.ctypes.data, v.ctypes.data, *args, img.ctypes.data, n) == 0
        W = np.random.RandomState(args[1]).randn(args[3] * args[5], args[4]).astype(np.float32)
        t = np.zeros(lib.gcnb_cheb_tap_image_bytes(args[3], args[4], args[5]), np.uint8)
        assert lib.gcnb_cheb_tap_image_build(W.ctypes.data, args[3], args[4], args[5], t.ctypes.data, t.size) == 0
        return img.tobytes(), t.tobytes()

    def fail(_):
        rc = lib.gcnb_cheb_tap_image_build(None, 32, 32, 5, None, 0)
        return rc, _lib.last_error()

    want = [build(j) for j in jobs]
    with ThreadPoolExecutor(8) as ex:
        futs = [ex.submit(build, jobs[i % 3]) for i in range(48)]
        bad = [ex.submit(fail, i) for i in range(8)]
        got = [f.result() for f in futs]
        errs = [f.result() for f in bad]
    assert all(g == want[i % 3] for i, g in enumerate(got))
    assert all(rc != 0 and "tap_image" in msg for rc, msg in errs)
