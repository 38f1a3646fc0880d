"""Host-side construction of the MoE design (`symphonypy/tl.py:368-392`).

The reference builds a dense one-hot matrix Phi_ [(B+1), N] with `pd.get_dummies` and a ridge
matrix diag(0, lamb...).  Phi is one-hot per batch variable, so the kernels take it as integer
level codes [V, N]; this module reproduces the reference's column order exactly: variables in
`key` order, levels of a variable sorted AS STRINGS (`obs[key].astype(str)` then get_dummies).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import _dist


def batch_codes(adata_query, key):
    """Returns (codes [V, N] int32, level names per variable)."""
    n = len(adata_query)
    if key is None:  # tl.py:368-369: one constant column "batch" = "1" -> a single level, code 0 everywhere
        return np.zeros((1, n), dtype=np.int32), _dist.union_levels([["1"]])
    if isinstance(key, str):
        cols = [np.asarray(adata_query.obs[key].astype(str), dtype=object)]
    else:
        cols = [np.asarray(adata_query.obs[k].astype(str), dtype=object) for k in key]
    local = [sorted(pd.unique(c).tolist()) for c in cols]
    levels = _dist.union_levels(local)
    codes = np.empty((len(cols), n), dtype=np.int32)
    for v, (c, lv) in enumerate(zip(cols, levels)):
        codes[v] = pd.Categorical(c, categories=lv).codes
    return codes, levels


def ridge_diagonal(lamb, n_levels):
    """diag of the (B+1)x(B+1) ridge matrix, entry 0 (intercept) = 0  (tl.py:380-392)."""
    phi_n = np.asarray(n_levels, dtype=int)
    if lamb is None:
        lamb = np.repeat([1] * len(phi_n), phi_n)
    elif isinstance(lamb, (float, int)):
        lamb = np.repeat([lamb] * len(phi_n), phi_n)
    elif len(lamb) == len(phi_n):
        lamb = np.repeat([lamb], phi_n)
    else:
        assert len(lamb) == np.sum(phi_n), "each batch variable must have a lambda"
    return np.insert(np.asarray(lamb, dtype=np.float64), 0, 0.0)
