"""Weighted-histogram and total-rate reductions on device arrays (np.histogram / np.sum semantics)."""
import numpy as np
import torch

from . import _native as N
from .runtime import Runtime


def histogram(x, weights, edges, out=None):
    """np.histogram(x, bins=edges, weights=weights)[0] for CUDA tensors: bins are half-open, the last one closed;
    values outside the edges (and NaN) are dropped.  `weights`: float64 or float32 tensor, or None (counts).
    Accumulates into `out` if given (so several flux components / GPUs can share one histogram)."""
    rt = Runtime.get(x.device)
    ed = edges if torch.is_tensor(edges) else rt.upload(np.asarray(edges, dtype=np.float64))
    n_edges = int(ed.shape[0])
    with rt.activate():
        if out is None:
            out = rt.zeros(n_edges - 1)
        w64 = weights if (weights is not None and weights.dtype == torch.float64) else None
        w32 = weights if (weights is not None and weights.dtype == torch.float32) else None
        if weights is not None and w64 is None and w32 is None:
            raise TypeError("weights must be float64 or float32")
        N.call("alpk_hist", N.ptr(x), N.ptr(w64), N.ptr(w32), int(x.shape[0]), N.ptr(ed), n_edges, N.ptr(out), rt.stream())
    return out


def total(x):
    """Deterministic two-stage sum of a float64 CUDA tensor -> 1-element tensor."""
    rt = Runtime.get(x.device)
    with rt.activate():
        out = rt.empty(1)
        N.call("alpk_sum", N.ptr(x), int(x.shape[0]), N.ptr(out), N.ptr(rt.scratch), rt.stream())
    return out
