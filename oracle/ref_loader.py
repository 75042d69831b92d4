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
