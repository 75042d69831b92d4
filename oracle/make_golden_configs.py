"""Golden vector of the reference's experiment configurations: runs ``gccn_model_common_param`` and the two network
builders of ``/root/reference/model.py`` (through ``oracle/ref_loader.load_model_builders``) over a grid of arguments
and stores what they return -- the model name and the hyper-parameter dictionary handed to ``cgcnn`` -- in
``tests/golden/ref_model_configs.json``.  Build container only (needs /root/reference); test infrastructure."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

CASES = {
    "common": [dict(modality="motor", training_samples=12800, target_name=["lf", "rf", "lh", "rh", "t"], block_dura=15),
               dict(modality="wm", training_samples=9000, target_name=["a", "b"], block_dura=1, eval_report=10, nepochs=20,
                    batch_size=64, hidden_size=128)],
    "chebyshev": [dict(Korder=5), dict(Korder=10), dict(Korder=3, flag_firstorder=1), dict(Korder=10, flag_firstorder=1),
                  dict()],
    "fourier": [dict(), dict(eigorders=0), dict(eigorders=0, dropout_lambda=0.25), dict(eigorders=5, dropout_lambda=0.5)],
}
LEVEL_SIZES = [372, 186]  # the pinned levels = 1 graph; the builders only read the shapes


def plain(v):
    if isinstance(v, np.ndarray):
        return [plain(x) for x in v.tolist()]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    return v


def main():
    ns = ref_loader.load_model_builders()
    L = [sp.identity(n, format="csr", dtype=np.float32) for n in LEVEL_SIZES]
    out = {"level_sizes": LEVEL_SIZES, "cases": []}
    with contextlib.redirect_stdout(io.StringIO()):
        for ci, ckw in enumerate(CASES["common"]):
            common = ns["gccn_model_common_param"](**ckw)
            out["cases"].append(dict(kind="common", common=ci, kwargs=ckw, params={k: plain(v) for k, v in common.items()}))
            for kind, fn in (("chebyshev", "build_chebyshev_graph_cnn"), ("fourier", "build_fourier_graph_cnn")):
                for kw in CASES[kind]:
                    model, name, params = ns[fn](dict(common), Laplacian_list=L, **kw)
                    assert {k: plain(v) for k, v in model.params.items()} == {k: plain(v) for k, v in params.items()}
                    out["cases"].append(dict(kind=kind, common=ci, kwargs=kw, name=name,
                                             params={k: plain(v) for k, v in params.items()}))
    path = os.path.join(ROOT, "tests", "golden", "ref_model_configs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
