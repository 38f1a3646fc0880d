"""world_size-2 gloo test of the multi-GPU host logic: query rows sharded across ranks,
level vocabulary agreed on by all ranks, MoE statistics summed over ranks BEFORE the priors are added,
then every rank solves redundantly.  The per-shard statistics are computed with numpy here (the CUDA
kernels produce the same arrays on a GPU); what is under test is `_dist` + `_design` + the ordering of
reduce / prior / solve."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import symphony_oracle as so


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_stats(Z, R, codes, n_levels):
    """gram [K,B1,B1], cross [K,B1,d] as `sym_moe_stats` leaves them (no intercept row, no priors)."""
    N, d = Z.shape
    K = R.shape[1]
    B1 = 1 + sum(n_levels)
    phi = np.zeros((B1, N))
    off = 1
    for v, lv in enumerate(n_levels):
        phi[off + codes[v], np.arange(N)] = 1
        off += lv
    gram = np.einsum("bn,nk,cn->kbc", phi, R, phi)
    cross = np.einsum("bn,nk,nd->kbd", phi, R, Z)
    return gram, cross


def _solve(gram, cross, Nr, C, lam, n_lev0):
    """numpy statement of `sym_moe_solve`."""
    K, B1, _ = gram.shape
    W = np.zeros((K, B1, cross.shape[2]))
    for k in range(K):
        A = gram[k].copy()
        diag = np.diag(A).copy()
        A[0, :] = diag
        A[:, 0] = diag
        A[0, 0] = diag[1:1 + n_lev0].sum() + Nr[k]
        H = cross[k].copy()
        H[0] = H[1:1 + n_lev0].sum(0) + C[k]
        W[k] = np.linalg.solve(A + np.diag(lam), H)
        W[k, 0] = 0
    return W


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pandas as pd

        from symphonypy_b200 import _design, _dist
        from symphonypy_b200.anndata_lite import AnnDataLite

        rng = np.random.default_rng(0)  # same stream on every rank: the full problem
        N, d, K = 600, 20, 12
        Z = rng.normal(size=(N, d))
        R = rng.dirichlet(np.ones(K), N)
        donors = rng.choice(["d1", "d10", "d2", "d3"], N)
        donors[: N // 2][donors[: N // 2] == "d3"] = "d1"   # level "d3" only exists on rank 1's shard
        chem = rng.choice(["v2", "v3"], N)
        Nr, C = rng.uniform(5, 20, K), rng.normal(size=(K, d))
        lo, hi = (0, N // 2) if rank == 0 else (N // 2, N)
        shard = AnnDataLite(np.zeros((hi - lo, 2), np.float32),
                            obs=pd.DataFrame({"donor": donors[lo:hi], "chem": chem[lo:hi]}))
        assert _dist.active() and _dist.world_size() == 2 and _dist.rank() == rank
        codes, levels = _design.batch_codes(shard, ["donor", "chem"])
        assert levels == [["d1", "d10", "d2", "d3"], ["v2", "v3"]], levels   # union, sorted as strings
        n_levels = [len(l) for l in levels]
        lam = _design.ridge_diagonal(None, n_levels)
        gram, cross = _shard_stats(Z[lo:hi], R[lo:hi], codes, n_levels)
        g, c = torch.from_numpy(gram), torch.from_numpy(cross)
        _dist.allreduce_stats(g, c)
        W = _solve(g.numpy(), c.numpy(), Nr, C, lam, n_levels[0])

        full = AnnDataLite(np.zeros((N, 2), np.float32), obs=pd.DataFrame({"donor": donors, "chem": chem}))
        phi_, lam_full = so.build_design(full, ["donor", "chem"], None)
        W_want = so.moe_weights(Z, phi_, R.T, Nr, C, lam_full)
        np.testing.assert_allclose(W, W_want, rtol=1e-9, atol=1e-10)
        # applying the shared weights shard by shard reproduces the full-data correction
        want = so.correct_query(Z, phi_, R.T, Nr, C, lam_full)
        mine = Z[lo:hi] - np.einsum("nk,kbd,bn->nd", R[lo:hi], W, phi_[:, lo:hi])
        np.testing.assert_allclose(mine, want[lo:hi], rtol=1e-9, atol=1e-10)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_moe(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_single_process_is_noop():
    from symphonypy_b200 import _dist

    g, c = torch.ones(2, 3, 3, dtype=torch.float64), torch.ones(2, 3, 4, dtype=torch.float64)
    _dist.allreduce_stats(g, c)
    assert not _dist.active() and float(g.sum()) == 18.0
    assert _dist.union_levels([["b", "a"]]) == [["b", "a"]]
