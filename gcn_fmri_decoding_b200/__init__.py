"""B200-native graph-convolution hot path of GCN_fmri_decoding."""
from . import graclus, graphs, synth  # noqa: F401  (host-side, NumPy/SciPy only)

__all__ = ["graclus", "graphs", "synth"]
