"""Import the reference's own host modules from /root/reference (build container only).

``lib_new/graph.py`` imports matplotlib at module level (used by ``plot_spectrum`` only,
``graph.py:3,131-143``); matplotlib is not installed, so an empty stub module is injected.
``lib_new/coarsening.py`` imports cleanly and runs its known-answer assert (``:217-218``) on
import.  ``lib_new/models_gcn.py`` cannot be imported (TensorFlow 1.x).

Used by ``oracle/make_golden.py`` and by the container-only tests; ``/root/reference`` does not
exist on the GPU box, where ``available()`` is False.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib_new", "graph.py"))


def load():
    """Return ``(graph, coarsening)`` modules of the reference."""
    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    if "matplotlib" not in sys.modules:
        stub = types.ModuleType("matplotlib")
        stub.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = stub
        sys.modules["matplotlib.pyplot"] = stub.pyplot
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from lib_new import coarsening, graph  # type: ignore
    return graph, coarsening


def load_model_builders():
    """The reference's configuration functions of ``model.py`` (``gccn_model_common_param :148-179``,
    ``build_fourier_graph_cnn :182-225``, ``build_chebyshev_graph_cnn :248-285``) executed as they are: the module cannot be
    imported (TensorFlow, keras, nilearn at the top), so the three function definitions are cut out of its AST and
    compiled in a namespace that holds what they read from their module -- NumPy, the globals of ``configure_fmri.py``
    (``atlas_name``, ``TR_step``) and a stand-in for ``models.cgcnn`` that records its arguments.  Returns the namespace."""
    import ast
    import types

    import numpy as np

    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    src = open(os.path.join(REFERENCE_ROOT, "model.py")).read()
    wanted = {"gccn_model_common_param", "build_fourier_graph_cnn", "build_chebyshev_graph_cnn"}
    tree = ast.parse(src)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {n.name for n in tree.body} == wanted
    cfg = {}
    exec(compile(open(os.path.join(REFERENCE_ROOT, "configure_fmri.py")).read(), "configure_fmri.py", "exec"), cfg)

    class RecordedModel:
        def __init__(self, config, L, **params):
            self.config, self.L, self.params = config, L, params

    ns = {"np": np, "atlas_name": cfg["atlas_name"], "TR_step": cfg["TR_step"], "config_TF": None,
          "models": types.SimpleNamespace(cgcnn=RecordedModel)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(tree, "model.py", "exec"), ns)
    return ns


def load_cgcnn_constructor():
    """``cgcnn.__init__`` of the reference (models_gcn.py:445-510) as a plain function ``init(obj, config, L, F, K, p, M,
    **kw)``: cut out of the class by AST (the module imports TensorFlow) and run on an object whose ``build_graph`` does
    nothing and whose layer methods exist as attributes -- what remains is the reference's argument checks and its
    Laplacian selection, executed from its own source."""
    import ast

    import numpy as np

    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "lib_new", "models_gcn.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "cgcnn")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
    # `super().__init__(config)` needs the class cell: drop that one statement (base_model.__init__ only stores config)
    fn.body = [s for s in fn.body if not (isinstance(s, ast.Expr) and isinstance(s.value, ast.Call)
                                          and "super" in ast.dump(s.value.func))]
    fn.name = "reference_cgcnn_init"
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np}
    exec(compile(mod, "models_gcn.py", "exec"), ns)

    class Shell:
        def build_graph(self, *a, **k):
            self.built = a

        chebyshev5 = chebyshev2 = fourier = spline = b1relu = b2relu = mpool1 = apool1 = staticmethod(lambda *a: None)

    return ns["reference_cgcnn_init"], Shell


def load_checkmat():
    """``BestCheckpointSaver`` and ``get_best_checkpoint`` of the reference's ``lib_new/checkmat.py`` (the module imports
    TensorFlow for its default saver only): the definitions are compiled from the reference source with ``tf`` left
    undefined -- callers pass their own ``saver``.  Returns the namespace."""
    import ast
    import glob
    import json

    import numpy as np

    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "lib_new", "checkmat.py")).read())
    tree.body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef))]
    ns = {"os": os, "glob": glob, "json": json, "np": np}
    exec(compile(tree, "checkmat.py", "exec"), ns)
    return ns


def load_base_model_method(name, extra=None):
    """One method of the reference's ``base_model`` (models_gcn.py:19-355) as a plain function taking ``self`` first,
    compiled from the reference source (the module itself imports TensorFlow).  Only methods whose body is host logic
    around ``sess.run`` make sense here (``predict``, ``evaluate``, ``fit``); the caller supplies an object with the
    attributes they touch and, through ``extra``, the module-level names the method reads (``checkmate`` for ``fit``)."""
    import ast
    import time

    import numpy as np
    import sklearn.metrics

    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "lib_new", "models_gcn.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "base_model")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    import collections
    import shutil

    ns = {"np": np, "time": time, "sklearn": __import__("sklearn"), "os": os, "sys": sys, "shutil": shutil,
          "collections": collections}
    ns["sklearn"].metrics = sklearn.metrics
    ns.update(extra or {})
    exec(compile(mod, "models_gcn.py", "exec"), ns)
    return ns[name]


def load_cgcnn_on_shim(variables, dtype=None, torch_autograd=False, dropout_masks=None):
    """The reference's ``base_model`` and ``cgcnn`` classes (models_gcn.py:17-682) compiled from their source with ``tf``
    bound to the NumPy stand-in of ``oracle/tf_shim.py`` and ``graph`` to the reference's own ``lib_new/graph.py``.
    ``build_graph`` is replaced by a no-op (it creates placeholders, the optimiser and the saver), everything else --
    ``__init__``, the layer methods, ``_inference``, ``loss``, the variable helpers -- is the reference's code.  Returns
    ``(cgcnn class, shim)``; ``variables`` maps TF variable names (``conv1/weights`` ...) to arrays.  With
    ``torch_autograd`` the stand-in works on torch tensors, so that the graph the reference code builds can be differentiated
    (``tf_shim.TorchShim``)."""
    import ast
    import collections
    import shutil
    import time

    import numpy as np
    import scipy.sparse
    import sklearn

    from oracle import tf_shim

    graph, _ = load()
    shim = tf_shim.TorchShim(variables, dropout_masks) if torch_autograd else tf_shim.Shim(variables, dtype or np.float32)
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "lib_new", "models_gcn.py")).read())
    tree.body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("base_model", "cgcnn")]
    ns = {"tf": shim, "graph": graph, "np": np, "scipy": scipy, "sklearn": sklearn, "os": os, "sys": sys, "time": time,
          "collections": collections, "shutil": shutil}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(tree, "models_gcn.py", "exec"), ns)
    ns["cgcnn"].build_graph = lambda self, *a, **k: None
    return ns["cgcnn"], shim


def load_model_perf(tf_stub):
    """The reference's ``model_perf`` class (models_gcn.py:936-1200) compiled from its source with ``tf`` bound to
    ``tf_stub`` (the caller's stand-in for the session / meta-graph calls of ``predict``).  Returns the class."""
    import ast
    from pathlib import Path

    import numpy as np
    import pandas as pd
    import sklearn
    import sklearn.metrics

    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "lib_new", "models_gcn.py")).read())
    tree.body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "model_perf"]
    ns = {"tf": tf_stub, "np": np, "pd": pd, "sklearn": sklearn, "os": os, "sys": sys, "Path": Path}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(tree, "models_gcn.py", "exec"), ns)
    return ns["model_perf"]
