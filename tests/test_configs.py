"""configs.py against the reference's own configuration functions (model.py:148-285): the golden file holds what the
reference functions returned in the build container (oracle/make_golden_configs.py); where /root/reference is present
they are executed live as well."""
import contextlib
import io
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN
from oracle import ref_loader


def _plain(v):
    if isinstance(v, np.ndarray):
        return [_plain(x) for x in v.tolist()]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


def _ours(golden):
    from gcn_fmri_decoding_b200 import configs

    L = [sp.identity(n, format="csr", dtype=np.float32) for n in golden["level_sizes"]]
    commons, out = {}, []
    with contextlib.redirect_stdout(io.StringIO()):
        for case in golden["cases"]:
            if case["kind"] == "common":
                commons[len(commons)] = configs.gccn_model_common_param(**case["kwargs"])
                out.append((None, commons[len(commons) - 1], None))
            else:
                fn = configs.build_chebyshev_graph_cnn if case["kind"] == "chebyshev" else configs.build_fourier_graph_cnn
                model, name, params = fn(dict(commons[case["common"]]), Laplacian_list=L, device="cpu", **case["kwargs"])
                out.append((name, params, model))
    return out


def test_builders_reproduce_the_reference_configurations():
    golden = json.load(open(os.path.join(GOLDEN, "ref_model_configs.json")))
    assert len(golden["cases"]) == 20
    for case, (name, params, model) in zip(golden["cases"], _ours(golden)):
        assert {k: _plain(v) for k, v in params.items()} == case["params"], (case["kind"], case["kwargs"])
        if case["kind"] == "common":
            continue
        assert name == case["name"]
        # the model was built with exactly these hyper-parameters
        p = case["params"]
        assert model.F == p["F"] and list(model.K) == p["K"] and model.p == p["p"] and model.M == p["M"]
        assert model.filter_name == p["filter"] and model.brelu_name == p["brelu"] and model.initial == p["initial"]
        assert (model.dropout, model.regularization, model.batch_size, model.num_epochs) == (
            p["dropout"], p["regularization"], p["batch_size"], p["num_epochs"])
        assert (model.learning_rate, model.decay_rate, model.momentum, model.eval_frequency, model.dir_name) == (
            p["learning_rate"], p["decay_rate"], p["momentum"], p["eval_frequency"], p["dir_name"])
        assert len(model.L) == 6 and all(l.shape[0] == golden["level_sizes"][0] for l in model.L)   # p = 1: level 0 throughout


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")
def test_golden_configurations_are_what_the_live_reference_returns():
    golden = json.load(open(os.path.join(GOLDEN, "ref_model_configs.json")))
    ns = ref_loader.load_model_builders()
    L = [sp.identity(n, format="csr", dtype=np.float32) for n in golden["level_sizes"]]
    commons = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for case in golden["cases"]:
            if case["kind"] == "common":
                commons[len(commons)] = ns["gccn_model_common_param"](**case["kwargs"])
                got = commons[len(commons) - 1]
            else:
                fn = "build_chebyshev_graph_cnn" if case["kind"] == "chebyshev" else "build_fourier_graph_cnn"
                _, name, got = ns[fn](dict(commons[case["common"]]), Laplacian_list=L, **case["kwargs"])
                assert name == case["name"]
            assert {k: _plain(v) for k, v in got.items()} == case["params"]


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")
def test_constructor_checks_and_laplacian_selection_against_the_reference_source(graph_l4):
    """cgcnn.__init__ (models_gcn.py:445-510) run from the reference's own source (TensorFlow parts cut away): the layers get
    the same Laplacians, the stored attributes agree, and what the reference asserts on, the replacement refuses."""
    from gcn_fmri_decoding_b200 import synth
    from gcn_fmri_decoding_b200.models import cgcnn

    init, Shell = ref_loader.load_cgcnn_constructor()
    _, _, _, L6 = synth.brain_graph(6)
    for L, p in ((graph_l4["L"], [4, 4]), (graph_l4["L"], [1] * 6), (L6, [1, 4, 1, 4, 1, 4]), (graph_l4["L"], [2, 2, 2]),
                 (graph_l4["L"], [8]), (graph_l4["L"], [2, 1, 4]), (graph_l4["L"], [1, 16])):
        n = len(p)
        kw = dict(F=[8] * n, K=[3] * n, p=p, M=[16, 5], filter="chebyshev5", brelu="b2relu", pool="mpool1", channel=15,
                  regularization=5e-4, dropout=0.5, batch_size=32, eval_frequency=7, dir_name="x")
        ref = Shell()
        with contextlib.redirect_stdout(io.StringIO()):
            init(ref, None, L, **kw)
        ours = cgcnn(None, L, device="cpu", **kw)
        assert [l.shape for l in ours.L] == [l.shape for l in ref.L]
        assert all((a != b).nnz == 0 for a, b in zip(ours.L, ref.L))
        assert ref.built == ((L[0].shape[0], 15),)                                # placeholder shape (M_0, channel)
        for attr in ("F", "K", "p", "M", "num_epochs", "learning_rate", "decay_rate", "decay_steps", "momentum",
                     "regularization", "dropout", "batch_size", "eval_frequency", "dir_name", "initial"):
            assert getattr(ours, attr) == getattr(ref, attr), attr
    bad = [dict(p=[3], F=[8], K=[3]),                      # not a power of two
           dict(p=[0], F=[8], K=[3]),                      # p >= 1
           dict(p=[8, 8], F=[8, 8], K=[3, 3])]             # needs 6 coarsening levels, 5 given
    for kw in bad:
        with pytest.raises(AssertionError), contextlib.redirect_stdout(io.StringIO()):
            init(Shell(), None, graph_l4["L"], M=[5], **kw)
        with pytest.raises(ValueError):
            cgcnn(None, graph_l4["L"], M=[5], device="cpu", **kw)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")
@pytest.mark.parametrize("maximize", [True, False])
def test_best_checkpoint_bookkeeping_against_the_reference_source(graph_l4, tmp_path, maximize):
    """checkpoints.BestCheckpoints against the reference's BestCheckpointSaver (checkmat.py:8-115, compiled from its source
    with a file-touching stand-in for tf.train.Saver) on random value sequences with ties: after every step both keep the
    same steps with the same values, and both name the same best checkpoint."""
    from gcn_fmri_decoding_b200 import checkpoints
    from gcn_fmri_decoding_b200.models import cgcnn

    ns = ref_loader.load_checkmat()

    class Saver:
        def save(self, sess, path, step):
            for ext in (".index", ".data-00000-of-00001", ".meta"):
                open("%s-%d%s" % (path, step, ext), "w").close()
            open(os.path.join(os.path.dirname(path), "checkpoint"), "w").close()

        def set_last_checkpoints_with_time(self, files):
            pass

    class Sess:
        def run(self, x):
            return x

    model = cgcnn(L=graph_l4["L"][4:], F=[4], K=[2], p=[1], M=[3], channel=2, device="cpu")
    rng = np.random.RandomState(5 + maximize)
    for trial in range(6):
        rdir, odir = str(tmp_path / ("ref%d" % trial)), str(tmp_path / ("ours%d" % trial))
        ref = ns["BestCheckpointSaver"](save_dir=rdir, num_to_keep=3, maximize=maximize, saver=Saver())
        ours = checkpoints.BestCheckpoints(odir, num_to_keep=3, maximize=maximize)
        values = rng.randint(0, 8, 14) / 8.0                     # few distinct values: ties happen
        for step, v in enumerate(values, start=1):
            ref.handle(v, Sess(), step * 10)
            ours.handle(v, model, step * 10)
            want = json.load(open(os.path.join(rdir, "best_checkpoints")))
            got = json.load(open(os.path.join(odir, "best_checkpoints")))
            assert {k + ".npz": x for k, x in want.items()} == got, (trial, step)
            kept_ref = sorted(f[:-len(".index")] for f in os.listdir(rdir) if f.endswith(".index"))
            kept_ours = sorted(f[:-len(".npz")] for f in os.listdir(odir) if f.endswith(".npz"))
            assert kept_ref == kept_ours == sorted(want)
        best_ref = ns["get_best_checkpoint"](rdir, select_maximum_value=maximize)
        if best_ref is not None:                                 # (the reference's function returns a path)
            assert os.path.basename(ours.best()) == os.path.basename(best_ref) + ".npz"


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")
def test_predict_batching_against_the_reference_source(graph_l4, monkeypatch):
    """cgcnn.predict against base_model.predict (models_gcn.py:31-71) compiled from the reference source and run on a
    stand-in session that evaluates the SAME stub network: zero-padded last batch, labels padded with zeros, the loss
    summed over batches and scaled by batch_size / size, a NaN batch loss counted as zero."""
    import torch

    from gcn_fmri_decoding_b200.models import cgcnn

    ref_predict = ref_loader.load_base_model_method("predict")
    m = cgcnn(L=graph_l4["L"], F=[32, 32], K=[5, 5], p=[4, 4], M=[512, 256, 22], channel=15, device="cpu", batch_size=8,
              regularization=0)
    rng = np.random.RandomState(3)
    proj = torch.tensor(rng.randn(15, 22).astype(np.float32))
    poison = {"batch": -1, "calls": 0}

    def net(x):                                             # stub network: logits from the first vertex's window
        lg = x[:, 0, :] @ proj
        poison["calls"] += 1
        if poison["calls"] - 1 == poison["batch"]:
            lg = lg * float("nan")
        return lg

    monkeypatch.setattr(m, "forward", lambda x, dropout=1.0, gather=None: net(x))

    class Sess:
        def run(self, fetches, feed):
            x = torch.tensor(np.asarray(feed["data"], np.float32))
            lg = net(x)
            pred = lg.argmax(1).numpy()
            if isinstance(fetches, list):
                lab = torch.tensor(np.asarray(feed["labels"]).astype(np.int64))
                return pred, float(m.loss(lg, lab))
            return pred

    class Ref:
        batch_size, ph_data, ph_labels, ph_dropout, op_prediction, op_loss = 8, "data", "labels", "dropout", "p", "l"

        def _get_session(self, sess):
            return Sess()

    data = rng.randn(21, 360, 15).astype(np.float32)
    labels = rng.randint(0, 22, 21)
    for bad_batch in (-1, 1):                               # without / with a NaN loss in the second batch
        poison.update(batch=bad_batch, calls=0)
        want_pred, want_loss = ref_predict(Ref(), data, labels)
        poison.update(calls=0)
        got_pred, got_loss = m.predict(data, labels)
        if bad_batch < 0:
            assert np.array_equal(got_pred, want_pred)
        assert abs(got_loss - want_loss) <= 1e-6 * abs(want_loss)
    poison.update(batch=-1, calls=0)
    assert np.array_equal(m.predict(data), ref_predict(Ref(), data))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")
def test_fit_and_evaluate_host_logic_against_the_reference_source(graph_l4, tmp_path, monkeypatch):
    """cgcnn.fit / evaluate against base_model.fit / evaluate (models_gcn.py:72-184) compiled from the reference source and
    run on stand-ins for the TensorFlow session and saver: with the same NumPy seed both draw the SAME batches in the
    same order, evaluate at the same steps, keep the same best checkpoints and return the same (accuracies, losses);
    evaluate returns the same summary string and figures."""
    import torch

    from gcn_fmri_decoding_b200.models import cgcnn

    ck = ref_loader.load_checkmat()
    import types

    ref_fit = ref_loader.load_base_model_method("fit", extra={"checkmate": types.SimpleNamespace(**ck)})
    ref_evaluate = ref_loader.load_base_model_method("evaluate")
    n, bs = 37, 8
    data = np.zeros((n, 360, 15), np.float32)
    data[:, 0, 0] = np.arange(n)                                   # window id, so that batches can be compared
    labels = np.arange(n) % 5
    val_scores = [52.0, 61.0, 58.0, 61.0, 70.0, 40.0, 66.0, 63.0, 59.0, 71.0, 30.0, 64.0]

    # ---- the reference, on stand-ins
    class Saver:
        def save(self, sess, path, step):
            step = sess.run(step) if isinstance(step, str) else step      # tf.train.Saver evaluates the step tensor
            open("%s-%d.index" % (path, step), "w").close()
            open(os.path.join(os.path.dirname(path), "checkpoint"), "w").close()

        def set_last_checkpoints_with_time(self, files):
            pass

    class Sess:
        def __init__(self):
            self.batches, self.step = [], 0

        def run(self, fetches, feed=None):
            if fetches == "init":
                return None
            if fetches == "global_step":
                return self.step
            ids = feed["data"][:, 0, 0].astype(int).tolist()
            assert feed["dropout"] == 0.5 and feed["labels"].tolist() == [i % 5 for i in ids]
            self.batches.append(ids)
            self.step += 1
            return 0.001, 1.0 / self.step

    class Ref:
        num_epochs, batch_size, eval_frequency, dropout = 3, bs, 2, 0.5
        ph_data, ph_labels, ph_dropout, op_train, op_loss_average, op_init = "data", "labels", "dropout", "t", "l", "init"
        global_step, op_saver, config, graph = "global_step", Saver(), None, None

        def __init__(self):
            self.sess = Sess()
            self.scores = iter(val_scores)

        def _get_session(self, sess):
            return sess

        def _get_path(self, folder):
            return str(tmp_path / "ref" / folder)

        def evaluate(self, d, l, sess=None, isTrain=False):
            return "accuracy: stub", next(self.scores), 0.0, 0.5

    ref = Ref()
    np.random.seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_acc, ref_losses, _ = ref_fit(ref, data, labels, data[:4], labels[:4])

    # ---- ours, with the step and the evaluation stubbed the same way
    m = cgcnn(L=graph_l4["L"], F=[32, 32], K=[5, 5], p=[4, 4], M=[512, 256, 22], channel=15, device="cpu",
              batch_size=bs, num_epochs=3, eval_frequency=2, dropout=0.5)
    seen, scores = [], iter(val_scores)

    class StubTrainer:
        def step(self, x, y, dropout=None):
            seen.append(x[:, 0, 0].long().tolist())
            return torch.tensor(1.0 / len(seen)), None

    monkeypatch.setattr(m, "evaluate", lambda d, l, **kw: ("accuracy: stub", next(scores), 0.0, 0.5))
    np.random.seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        acc, losses, _ = m.fit(data, labels, data[:4], labels[:4], best_checkpoint_dir=str(tmp_path / "ours"),
                               trainer=StubTrainer())
    assert seen == ref.sess.batches and len(seen) == int(3 * n / bs)      # the same windows in the same order
    assert acc == ref_acc and losses == ref_losses
    want = json.load(open(tmp_path / "ref" / "checkpoints" / "model" / "best_checkpoints"))
    got = json.load(open(tmp_path / "ours" / "best_checkpoints"))
    assert {k + ".npz": v for k, v in want.items()} == got

    # ---- evaluate: summary string and figures
    rng = np.random.RandomState(2)
    y = rng.randint(0, 5, 60)
    pred = np.where(rng.rand(60) < 0.7, y, rng.randint(0, 5, 60)).astype(np.float64)

    class RefEval:
        sess = object()

        def _get_session(self, sess):
            return sess

        def predict(self, d, l, sess):
            return pred, 0.4321

    want = ref_evaluate(RefEval(), None, y, isTrain=True)
    monkeypatch.undo()
    monkeypatch.setattr(m, "predict", lambda d, l=None, **kw: (pred, 0.4321))
    got = m.evaluate(None, y)
    assert got[0] == want[0] and got[1:] == pytest.approx(want[1:], rel=1e-12)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present (GPU box)")
def test_model_perf_predict_against_the_reference_source(graph_l4, tmp_path, monkeypatch):
    """perf.model_perf.predict against the reference's model_perf.predict (models_gcn.py:960-1088) compiled from its source,
    its TensorFlow session replaced by a stand-in that evaluates the same stub network: predicted labels, the reported
    loss (the reference SUMS the batch losses here, unlike base_model.predict), the accuracy figure, the per-subject F1
    table and the per-time-point accuracy table."""
    import types

    import torch

    from gcn_fmri_decoding_b200 import tf_bundle
    from gcn_fmri_decoding_b200.models import cgcnn
    from gcn_fmri_decoding_b200.perf import model_perf

    rng = np.random.RandomState(8)
    C, bs, n_sub, per_sub, dura = 6, 16, 5, 34, 17
    n = n_sub * per_sub
    proj = torch.tensor(rng.randn(15, C).astype(np.float32))
    m = cgcnn(L=graph_l4["L"], F=[32, 32], K=[5, 5], p=[4, 4], M=[64, 32, C], channel=15, device="cpu", regularization=0)

    def net(x):
        return x[:, 0, :] @ proj

    monkeypatch.setattr(m, "forward", lambda x, dropout=1.0, gather=None: net(x))

    class Sess:
        graph = types.SimpleNamespace(get_operations=lambda: [], get_tensor_by_name=lambda name: name)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def run(self, fetches, feed):
            lg = net(torch.tensor(np.asarray(feed["inputs/data:0"], np.float32)))
            lab = torch.tensor(np.asarray(feed["inputs/labels:0"]).astype(np.int64))
            return lg.numpy(), lg.argmax(1).numpy(), float(m.loss(lg, lab))

    tf_stub = types.SimpleNamespace(
        reset_default_graph=lambda: None, Session=Sess,
        train=types.SimpleNamespace(import_meta_graph=lambda path, clear_devices=True: types.SimpleNamespace(
            restore=lambda sess, path: None)))
    RefPerf = ref_loader.load_model_perf(tf_stub)

    run = tmp_path / "run"
    tf_bundle.save_tf_checkpoint(m, str(run / "model" / "best.ckpt"), step=40)
    (run / "model" / "checkpoint").write_text('model_checkpoint_path: "best.ckpt-40"\nall_model_checkpoint_paths: "best.ckpt-40"\n')
    (tmp_path / "train_logs").mkdir()
    monkeypatch.chdir(tmp_path)                      # the reference writes train_logs/<...>.csv relative to the cwd
    # labels cycle so that every subject and every time point sees every class (sklearn refuses empty groups); the windows
    # are built so that the stub network predicts the label 70 % of the time
    labels = (np.arange(n) // dura + np.arange(n) % dura) % C
    wanted = np.where(rng.rand(n) < 0.7, labels, rng.randint(0, C, n))
    data = (rng.randn(n, 360, 15) * 0.01).astype(np.float32)
    data[:, 0, :] += (5 * np.eye(C)[wanted] @ np.linalg.pinv(proj.numpy())).astype(np.float32)
    assert np.array_equal(net(torch.tensor(data)).argmax(1).numpy(), wanted)
    names = ["task_%d_x" % i for i in range(C)]
    subs = ["s%d" % i for i in range(n_sub)]
    for kw in (dict(), dict(sub_name=subs), dict(trial_dura=dura, flag_starttr=True)):
        with contextlib.redirect_stdout(io.StringIO()):
            _, want_lab, want_loss, want_acc = RefPerf().predict(str(run), data, labels, target_name=names, batch_size=bs, **kw)
            got_lg, got_lab, got_loss, got_acc = model_perf().predict(str(run), data, labels, target_name=names,
                                                                      batch_size=bs, model=m, **kw)
        assert np.array_equal(got_lab, want_lab)
        assert abs(got_loss - want_loss) <= 1e-6 * abs(want_loss)
        assert np.allclose(np.asarray(got_acc, np.float64), np.asarray(want_acc, np.float64), rtol=0, atol=1e-9), kw
        assert got_lg.shape == (n, C) and np.array_equal(got_lg.argmax(1), want_lab)
