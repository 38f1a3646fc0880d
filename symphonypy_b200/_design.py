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


def _column_codes(col: pd.Series):
    """(local sorted level names as strings, per-cell code into them) of one obs column, equal to what
    `obs[key].astype(str)` + `pd.get_dummies` give (`tl.py:370-376`) — without converting N cells to Python strings:
    only the DISTINCT values are stringified (a categorical column's categories, or the uniques of a factorize)."""
    if isinstance(col.dtype, pd.CategoricalDtype):
        raw = col.cat.codes.to_numpy()
        uniq = col.cat.categories
        if (raw < 0).any():          # astype(str) turns a missing value into the string "nan"
            uniq = uniq.astype(object).append(pd.Index([np.nan], dtype=object))
            raw = np.where(raw < 0, len(uniq) - 1, raw)
    else:
        raw, uniq = pd.factorize(col, use_na_sentinel=False)
    # str() of the DISTINCT values only.  A missing value becomes the level "nan", as `astype(str)` made it
    # when the reference was written (pandas < 3; pandas >= 3 keeps it missing and get_dummies then drops the cell
    # from every level of that variable — a cell without a batch is not something the MoE model defines).
    names = np.asarray([str(v) for v in np.asarray(uniq, dtype=object)], dtype=object)
    used = np.zeros(len(names), dtype=bool)
    used[raw] = True
    # distinct values may collide after str() (1 and "1"); unused categories give no dummy column
    levels = sorted(set(names[used].tolist()))
    lut = np.full(len(names), -1, dtype=np.int32)
    pos = {lv: i for i, lv in enumerate(levels)}
    for i in np.nonzero(used)[0]:
        lut[i] = pos[names[i]]
    return levels, lut, raw


def batch_codes(adata_query, key):
    """Returns (codes [V, N] int32, level names per variable)."""
    n = len(adata_query)
    if key is None:  # tl.py:368-369: one constant column "batch" = "1" -> a single level, code 0 everywhere
        return np.zeros((1, n), dtype=np.int32), _dist.union_levels([["1"]])
    keys = [key] if isinstance(key, str) else list(key)
    parts = [_column_codes(adata_query.obs[k]) for k in keys]
    local = [p[0] for p in parts]
    levels = _dist.union_levels(local)        # under torch.distributed: the sorted union over all shards
    codes = np.empty((len(keys), n), dtype=np.int32)
    for v, ((lv_local, lut, raw), lv) in enumerate(zip(parts, levels)):
        if lv != lv_local:                    # re-base the local codes on the agreed vocabulary
            where = {name: i for i, name in enumerate(lv)}
            remap = np.asarray([where[name] for name in lv_local], dtype=np.int32)
            lut = np.where(lut >= 0, remap[np.clip(lut, 0, None)], -1).astype(np.int32)
        codes[v] = lut[raw]
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
