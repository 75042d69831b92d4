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
