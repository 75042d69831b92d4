"""Checkpoints under the reference's variable names, and the "best k" bookkeeping of its training loop.

The reference saves TensorFlow-1.x checkpoints (``tf.train.Saver``, ``lib_new/models_gcn.py:220``) whose variables are
named ``conv{i}/weights``, ``conv{i}/bias``, ``fc{i}/weights`` ... ``logits/bias`` (``:662``, ``:343``, ``:351``, ``:675``,
``:680``); ``fit`` keeps the best three by validation accuracy (``lib_new/checkmat.py:8-84``, used at
``models_gcn.py:127,175``) and ``evaluate`` restores the latest one (``:89-90``).  TensorFlow is not available here, so
the native container is a NumPy ``.npz`` with exactly those names (``W`` as ``[Fin*K, Fout]``, row ``f*K + k``): a
reference-side ``{v.name[:-2]: sess.run(v) for v in tf.trainable_variables()}`` dump loads unchanged.  TensorFlow's own
V2 checkpoint files (``<prefix>.index`` / ``.data-*``) are read and written by ``tf_bundle.py`` in plain Python;
``load_checkpoint`` accepts such a prefix too.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

INDEX_NAME = "best_checkpoints"  # the JSON index the reference keeps in the checkpoint directory


def save_checkpoint(model, path, step=None):
    """Write the model's parameters as ``<path>[-<step>].npz`` under the TF variable names; returns the file name."""
    name = path if step is None else "%s-%d" % (path, int(step))
    if not name.endswith(".npz"):
        name += ".npz"
    os.makedirs(os.path.dirname(os.path.abspath(name)), exist_ok=True)
    np.savez(name, **{k.replace("/", "__"): v for k, v in model.state_dict_tf().items()})
    return name


def load_checkpoint(model, path):
    """Load a file written by ``save_checkpoint`` (or any ``.npz`` keyed by the TF variable names, ``/`` or ``__``), or a
    TensorFlow V2 checkpoint given by its prefix (``<path>.index`` exists: read by ``tf_bundle``, no TensorFlow needed)."""
    if os.path.exists(path + ".index"):
        from . import tf_bundle

        return tf_bundle.load_tf_checkpoint(model, path)
    with np.load(path) as z:
        d = {k.replace("__", "/"): z[k] for k in z.files}
    model.load_state_dict_tf(d)
    return model


def latest_checkpoint(directory, prefix="model"):
    """The checkpoint with the highest step in ``directory`` (``tf.train.latest_checkpoint``), or None: a ``.npz`` file
    of ``save_checkpoint``, or the prefix of a TensorFlow checkpoint (``<prefix>-<step>.index``) -- either loads with
    ``load_checkpoint``."""
    best, best_step = None, -1
    for ext in (".npz", ".index"):
        for f in glob.glob(os.path.join(directory, prefix + "-*" + ext)):
            try:
                step = int(os.path.basename(f)[len(prefix) + 1:-len(ext)])
            except ValueError:
                continue
            if step > best_step:
                best, best_step = (f if ext == ".npz" else f[:-len(ext)]), step
    return best


class BestCheckpoints:
    """Keep the ``num_to_keep`` best checkpoints of a run in ``save_dir``, ranked by a validation value.

    ``handle(value, model, step)`` has the semantics of the reference's ``BestCheckpointSaver.handle``: below capacity
    every checkpoint is kept; at capacity a new one replaces the worst unless every kept value is at least as good.
    The JSON index ``best_checkpoints`` maps file names to values, as in the reference.
    """

    def __init__(self, save_dir, num_to_keep=3, maximize=True, prefix="best.ckpt"):
        self.save_dir, self.num_to_keep, self.maximize, self.prefix = save_dir, int(num_to_keep), bool(maximize), prefix
        os.makedirs(save_dir, exist_ok=True)
        self.index_file = os.path.join(save_dir, INDEX_NAME)

    def _load(self):
        if not os.path.exists(self.index_file):
            return {}
        with open(self.index_file) as f:
            return json.load(f)

    def _store(self, index):
        with open(self.index_file, "w") as f:
            json.dump(index, f, indent=3)

    def _ranked(self, index):
        return sorted(index, key=index.get, reverse=self.maximize)  # best first

    def handle(self, value, model, step):
        """Returns the file written, or None when the checkpoint did not make it into the best ``num_to_keep``."""
        value = float(value)
        index = self._load()
        name = "%s-%d.npz" % (self.prefix, int(step))
        if len(index) >= self.num_to_keep:
            better = (lambda kept: kept >= value) if self.maximize else (lambda kept: kept <= value)
            if all(better(v) for v in index.values()):
                return None
            worst = self._ranked(index)[-1]
            index.pop(worst)
            try:
                os.remove(os.path.join(self.save_dir, worst))
            except FileNotFoundError:
                pass
        index[name] = value
        self._store(index)
        return save_checkpoint(model, os.path.join(self.save_dir, name))

    def best(self):
        """Path of the best checkpoint kept so far (``checkmat.get_best_checkpoint``), or None."""
        index = self._load()
        return os.path.join(self.save_dir, self._ranked(index)[0]) if index else None


def classification_summary(labels, predictions, loss=None):
    """The figures ``cgcnn.evaluate`` reports (``models_gcn.py:103-109``): accuracy and weighted F1 in percent, the
    number of correct predictions and the reference's one-line summary string."""
    labels, predictions = np.asarray(labels).astype(np.int64), np.asarray(predictions).astype(np.int64)
    n = len(labels)
    ncorrect = int((labels == predictions).sum())
    accuracy = 100.0 * ncorrect / max(n, 1)
    # weighted F1: per-class F1 averaged with the class supports (sklearn.metrics.f1_score(average='weighted'))
    f1 = 0.0
    for c in np.unique(labels):
        tp = float(((predictions == c) & (labels == c)).sum())
        fp = float(((predictions == c) & (labels != c)).sum())
        fn = float(((predictions != c) & (labels == c)).sum())
        denom = 2 * tp + fp + fn
        f1 += (2 * tp / denom if denom > 0 else 0.0) * float((labels == c).sum())
    f1 = 100.0 * f1 / max(n, 1)
    string = "accuracy: {:.2f} ({:d} / {:d}), f1 (weighted): {:.2f}".format(accuracy, ncorrect, n, f1)
    if loss is not None:
        string += ", loss: {:.2e}".format(loss)
    return string, accuracy, f1


def confusion_matrix(labels, predictions, n_classes):
    """``C[i, j]`` = windows of class ``i`` predicted as ``j`` over ``range(n_classes)`` -- what the reference prints with
    ``sklearn.metrics.confusion_matrix(labels, predictions, labels=range(len(target_name)))`` (models_gcn.py:98-99);
    labels or predictions outside the range are left out, as sklearn does."""
    labels, predictions = np.asarray(labels).astype(np.int64), np.asarray(predictions).astype(np.int64)
    ok = (labels >= 0) & (labels < n_classes) & (predictions >= 0) & (predictions < n_classes)
    C = np.zeros((n_classes, n_classes), np.int64)
    np.add.at(C, (labels[ok], predictions[ok]), 1)
    return C


def classification_report(labels, predictions, target_names):
    """Per-class precision / recall / F1 / support table over ``range(len(target_names))`` plus the accuracy, macro and
    weighted averages -- the text the reference prints through ``sklearn.metrics.classification_report`` when
    ``evaluate`` is given ``target_name`` (models_gcn.py:94-97).  Returns ``(text, rows)`` with ``rows[name] =
    (precision, recall, f1, support)``; undefined ratios count as 0, like sklearn's ``zero_division`` default."""
    labels, predictions = np.asarray(labels).astype(np.int64), np.asarray(predictions).astype(np.int64)
    n_classes = len(target_names)
    rows, width = {}, max([len(str(t)) for t in target_names] + [len("weighted avg")])
    head = "{:>{w}s} {:>9s} {:>9s} {:>9s} {:>9s}".format("", "precision", "recall", "f1-score", "support", w=width)
    lines = [head, ""]
    P, R, F, S = [], [], [], []
    for c, name in enumerate(target_names):
        tp = float(((predictions == c) & (labels == c)).sum())
        npred, support = float((predictions == c).sum()), int((labels == c).sum())
        prec = tp / npred if npred > 0 else 0.0
        rec = tp / support if support > 0 else 0.0
        f1 = 2 * prec * rec / (prec + rec) if prec + rec > 0 else 0.0
        rows[str(name)] = (prec, rec, f1, support)
        P.append(prec), R.append(rec), F.append(f1), S.append(support)
        lines.append("{:>{w}s} {:>9.2f} {:>9.2f} {:>9.2f} {:>9d}".format(str(name), prec, rec, f1, support, w=width))
    total = int(np.sum(S))
    Wt = np.asarray(S, np.float64) / max(total, 1)
    # sklearn prints 'accuracy' when the label set covers every label seen, 'micro avg' otherwise; with all classes
    # listed the two coincide
    acc = float((labels == predictions).sum()) / max(len(labels), 1)
    rows["accuracy"] = (acc, acc, acc, total)
    rows["macro avg"] = (float(np.mean(P)), float(np.mean(R)), float(np.mean(F)), total)
    rows["weighted avg"] = (float(np.dot(P, Wt)), float(np.dot(R, Wt)), float(np.dot(F, Wt)), total)
    lines.append("")
    lines.append("{:>{w}s} {:>9s} {:>9s} {:>9.2f} {:>9d}".format("accuracy", "", "", acc, total, w=width))
    for key in ("macro avg", "weighted avg"):
        lines.append("{:>{w}s} {:>9.2f} {:>9.2f} {:>9.2f} {:>9d}".format(key, *rows[key], w=width))
    return "\n".join(lines) + "\n", rows
