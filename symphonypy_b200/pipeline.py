"""The device-resident hot path as one function: assign -> MoE (stats, all-reduce, solve, apply).

`tl.map_embedding` calls this after the projection; `bench.py` calls it with inputs already in
HBM.  `mark(name)` (optional) is invoked after each stage so a caller can drop CUDA events.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _dist, engine


def soft_assign_and_correct(Z, Y, inv_sigma, codes, n_levels, lam, Nr, C, mark=None, R=None):
    """Z [N,d] f32 (device) -> (R [N,K], Zcorr [N,d]).   `symphonypy/tl.py:358-397`.

    codes [V,N] int32 level codes, n_levels per variable, lam [B+1] f64 ridge diagonal,
    Nr [K] / C [K,d] f64 reference priors (device).  Raises LinAlgError when a ridge system is singular.
    `R`: soft assignments already computed by the caller (streamed per chunk behind the projection), else None.
    """
    mark = mark or (lambda name: None)
    if R is None:
        R = engine.assign(Z, Y, inv_sigma)
    mark("assign")
    orders = [engine.batch_order(codes[v], n_levels[v]) for v in range(len(n_levels))]
    gram, cross = engine.moe_stats(Z, R, codes, n_levels, orders)
    mark("moe_stats")
    _dist.allreduce_stats(gram, cross)  # the only cross-GPU exchange of the path (priors are added after it)
    mark("allreduce")
    W, info = engine.moe_solve(gram, cross, Nr, C, lam, n_levels[0])
    Zc = engine.moe_apply(Z, R, codes, n_levels, orders, W)
    mark("moe_solve_apply")
    return R, Zc, info


def check_solve(info: torch.Tensor) -> None:
    if int(info.abs().max().item()) != 0:
        raise np.linalg.LinAlgError("Singular matrix")  # what `np.linalg.inv` raises at `_utils.py:208`


def knn_transfer(index: engine.KnnIndex, Q, k: int, ref_codes, method=engine.KNN_AUTO, mark=None, events=None,
                 prune: bool = False):
    """Q [N,d] -> (label codes [N], vote fraction [N], idx [N,k], dist2 [N,k], repaired rows)."""
    mark = mark or (lambda name: None)
    idx, dist2, n_rep = engine.knn_search_certified(index, Q, k, method, events=events, prune=prune)
    mark("knn")
    lab, conf = engine.vote(idx, ref_codes)
    mark("vote")
    return lab, conf, idx, dist2, n_rep
