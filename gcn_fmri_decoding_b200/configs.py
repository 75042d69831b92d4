"""The reference's experiment configurations: ``gccn_model_common_param`` and the model builders
of ``/root/reference/model.py`` (``:148-179`` common settings, ``:182-225`` spectral network, ``:248-285`` ChebyNet /
first-order network) with the same names, arguments, hyper-parameter values and ``(model, name, params)`` return value --
``predict_states.py:59-66`` and ``training.py`` build their networks through exactly these calls.  The only additions are
keyword arguments that tell ``cgcnn`` where to live (``device``) and how raw windows are permuted (``perm``,
``n_input_vertices``); the values of ``configure_fmri.py`` the reference reads as globals (``atlas_name = 'MMP'``,
``TR_step = 1``) are keyword arguments with those defaults.  The spline network (``:228-246``) is not offered: the
``spline`` filter is outside the scope of this repository.
"""
from __future__ import annotations

import numpy as np

from .models import cgcnn


def gccn_model_common_param(modality, training_samples, target_name=None, block_dura=15, eval_report=20, nepochs=100,
                            batch_size=128, layers=6, pool_size=1, hidden_size=256, atlas_name="MMP", TR_step=1):
    """Settings shared by all networks of an experiment (model.py:148-179): ``C = len(target_name) + 1`` classes, 5e-4 L2,
    dropout keep 0.5, b2relu + mpool1, ``eval_report`` evaluations over the run.  ``layers`` / ``pool_size`` are accepted
    and unused, as in the reference."""
    C = len(target_name) + 1
    dir_name = "%s/%s_win%d/c%d" % (modality, atlas_name, block_dura, len(target_name))
    if TR_step > 1:
        dir_name += "_step%d/" % TR_step
    decay_steps = training_samples / batch_size
    return {
        "dir_name": dir_name, "num_epochs": nepochs, "batch_size": batch_size, "decay_steps": decay_steps,
        "eval_frequency": int(nepochs * decay_steps / eval_report),
        "brelu": "b2relu", "pool": "mpool1", "initial": "normal",
        "regularization": 5e-4, "dropout": 0.5, "learning_rate": 0.05, "decay_rate": 0.95, "momentum": 0.9,
        "channel": block_dura,
        "F": [32, 32, 64, 64, 128, 128, 128, 128], "K": [20, 10, 10, 10, 5, 5, 5, 5], "p": [1, 4, 1, 4, 1, 4, 1, 1],
        "M": [hidden_size, C],
    }


def _announce(kind, Laplacian_list):
    print("\nBuilding convolutional layers with %s\n" % kind)
    if not Laplacian_list:
        print("Laplacian matrix for multi-scale graphs are requried!")
    else:
        print("Laplacian matrix for multi-scale graphs:")
        print([l.shape for l in Laplacian_list])


def build_chebyshev_graph_cnn(gcnn_common, Laplacian_list=None, Korder=5, flag_firstorder=0, **where):
    """ChebyNet (six 32-filter layers of order ``Korder``, no pooling) or, with ``flag_firstorder``, the first-order
    network (``K = 1``) -- model.py:248-285: Adam-era settings ``learning_rate`` 0.001, ``decay_rate`` 0.9, He
    initialisation, head ``[2 * hidden, hidden, C]``.  ``where``: ``device=``, ``perm=``, ``n_input_vertices=``, ``seed=``."""
    _announce("Chebyshev polynomial", Laplacian_list)
    name = "cgconv_cgconv_fc_softmax_K%d" % Korder
    if flag_firstorder:
        name = "cgconv_cgconv_fc_softmax_firstorder"
    if Korder == 10:
        name = "cgconv_cgconv_fc_softmax"
    params = dict(gcnn_common)
    params["dir_name"] += name
    params.update(filter="chebyshev5", learning_rate=0.001, decay_rate=0.9, initial="he", F=[32] * 6, p=[1] * 6,
                  K=[1] * 6 if flag_firstorder else [Korder] * 6, M=[gcnn_common["M"][0] * 2] + list(gcnn_common["M"]))
    print(params)
    model = cgcnn(None, Laplacian_list, **params, **where)
    return model, name, params


def build_fourier_graph_cnn(gcnn_common, Laplacian_list=None, dropout_lambda=0.0, eigorders=10, **where):
    """Spectral network (six 32-filter ``fourier`` layers, no pooling) -- model.py:182-225.  ``K`` is carried in the
    parameters as the reference does (``eigorders`` per layer, or with ``eigorders`` false the vertex count of the level
    each layer works on, scaled by ``1 - dropout_lambda``); the ``fourier`` filter itself ignores it (models_gcn.py:530)."""
    _announce("fourier basis of Laplacian", Laplacian_list)
    name = "_".join(("fgconv_fgconv_fc_softmax", "K%d" % eigorders if eigorders else "full",
                     "drop%s" % dropout_lambda if dropout_lambda > 0 else ""))
    if eigorders == 10:
        name = "fgconv_fgconv_fc_softmax"
    params = dict(gcnn_common)
    params["dir_name"] += name
    params.update(filter="fourier", F=[32] * 6, p=[1] * 6, M=[gcnn_common["M"][0] * 2] + list(gcnn_common["M"]))
    if eigorders:
        params["K"] = [eigorders] * 6
    else:
        sizes = [l.shape[0] for l in Laplacian_list]
        K = np.zeros(len(params["p"]), dtype=int)
        K[0] = int(sizes[0] * (1 - dropout_lambda))
        level = 0
        for li, pi in enumerate(params["p"][:-1]):
            if pi == 1:
                K[li + 1] = K[li]
            else:  # pooling by 2 / 4 moves one / two coarsening levels down
                level += 1 if pi == 2 else 2
                K[li + 1] = int(sizes[level] * (1 - dropout_lambda))
        params["K"] = K
        print(params["K"])
    model = cgcnn(None, Laplacian_list, **params, **where)
    return model, name, params
