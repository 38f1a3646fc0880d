"""Seeded synthetic single-cell data of the shapes BASELINE.json names.

There is no network for datasets, so tests, `bench.py` and `smoke()` all use
this generator.  It is written in torch so the same code fills
a CPU tensor for the small parity cases and generates straight into HBM for the
full-size configurations (config 3's expression matrix is 80 GB and never
exists on the host).

Generative model
  * L cell types, each with a "program" (10 % of the genes) whose detection
    probability is raised from 5 % to 35 %  -> overall density ~8 %, values
    Gamma(2, 0.8) (log1p-like, non-negative, float32).
  * Query cells additionally carry V batch variables; level b of variable v
    scales a level-specific 5 % gene subset by 1.5 (the effect the MoE
    correction removes).
  * Reference: gene mean / std estimated from reference cells, loadings U =
    orthonormalised Gaussian [G, d], embedding = clip((X-mu)/sd, +-10) @ U,
    labels = cell type, Harmony summary (C, Nr, K) from a soft assignment of the
    reference embedding to K perturbed reference cells — the exact fields
    `symphonypy/_utils.py:44-61` stores and `symphonypy/tl.py:346-395` reads.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


@dataclass
class SynthModel:
    """Everything shared between reference and query generation."""

    G: int
    d: int
    L: int
    device: torch.device
    programs: torch.Tensor  # [L, G] bool
    mean: torch.Tensor = None  # [G] f32
    std: torch.Tensor = None  # [G] f32
    U: torch.Tensor = None  # [G, d] f32
    batch_sets: dict = field(default_factory=dict)

    def expr(self, comp: torch.Tensor, gen: torch.Generator, codes: torch.Tensor | None = None,
             seed: int = 0) -> torch.Tensor:
        """Dense [n, G] float32 expression for cells of the given types."""
        n = comp.shape[0]
        dev = self.device
        prob = torch.where(self.programs[comp], 0.35, 0.05)
        u = torch.rand((n, self.G), generator=gen, device=dev)
        mask = u < prob
        del u, prob
        e = torch.rand((n, self.G), generator=gen, device=dev).clamp_min_(1e-12).log_().neg_()
        e2 = torch.rand((n, self.G), generator=gen, device=dev).clamp_min_(1e-12).log_().neg_()
        x = (e + e2).mul_(0.8)
        del e, e2
        x = torch.where(mask, x, torch.zeros((), device=dev))
        if codes is not None:
            for v in range(codes.shape[0]):
                bs = self.batch_set(v, int(codes[v].max().item()) + 1, seed)
                x = torch.where(bs[codes[v].long()], x * 1.5, x)
        return x.float()

    def batch_set(self, v: int, levels: int, seed: int) -> torch.Tensor:
        key = (v, levels, seed)
        if key not in self.batch_sets:
            g = _gen("cpu", 7919 * (v + 1) + seed)
            self.batch_sets[key] = (torch.rand((levels, self.G), generator=g) < 0.05).to(self.device)
        return self.batch_sets[key]

    def project(self, x: torch.Tensor) -> torch.Tensor:
        """Plain torch statement of `_utils.py:298-308` (setup only, never timed)."""
        t = ((x - self.mean) / self.std).clamp_(-10.0, 10.0)
        return t @ self.U


def make_model(G: int, d: int, L: int, seed: int, device) -> SynthModel:
    device = torch.device(device)
    g = _gen("cpu", seed)
    programs = (torch.rand((L, G), generator=g) < 0.10).to(device)
    q, _ = torch.linalg.qr(torch.randn((G, d), generator=g, dtype=torch.float64))
    m = SynthModel(G=G, d=d, L=L, device=device, programs=programs)
    m.U = q.float().contiguous().to(device)
    # gene statistics from a fixed sample of reference-like cells
    gs = _gen(device, seed + 1)
    comp = torch.randint(0, L, (4096,), generator=gs, device=device)
    x = m.expr(comp, gs)
    m.mean = x.mean(0)
    m.std = x.std(0).clamp_min_(0.05)
    return m


def make_reference(model: SynthModel, M: int, K: int, seed: int, chunk: int = 32768, keep_R: bool = False):
    """Reference embedding [M, d], type codes [M], and the Harmony summary (plus the reference's own soft
    assignment R_ref [M, K] when `keep_R`: `uns["harmony"]["R"]`, needed by `per_cell_confidence`)."""
    dev = model.device
    g = _gen(dev, seed + 2)
    Z = torch.empty((M, model.d), device=dev, dtype=torch.float32)
    comp = torch.randint(0, model.L, (M,), generator=g, device=dev)
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        Z[s:e] = model.project(model.expr(comp[s:e], g))
    # K centroids: perturbed reference cells; soft assignment as in `_utils.py:168-176`
    pick = torch.randperm(M, generator=_gen("cpu", seed + 3))[:K].to(dev)
    cent = Z[pick] + 0.05 * torch.randn((K, model.d), generator=g, device=dev)
    Y = cent / cent.norm(dim=1, keepdim=True)
    Nr = torch.zeros(K, device=dev, dtype=torch.float64)
    C = torch.zeros((K, model.d), device=dev, dtype=torch.float64)
    R_all = torch.empty((M, K), device=dev, dtype=torch.float64) if keep_R else None
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        z = Z[s:e]
        zc = z / z.max(dim=1, keepdim=True).values
        zc = zc / zc.norm(dim=1, keepdim=True)
        r = torch.softmax(-2.0 * (1.0 - zc @ Y.T) / 0.1, dim=1).double()
        Nr += r.sum(0)
        C += r.T @ z.double()
        if keep_R:
            R_all[s:e] = r
    if keep_R:
        return Z, comp, C, Nr, R_all
    return Z, comp, C, Nr


def make_query(model: SynthModel, N: int, levels: list[int], seed: int, chunk: int = 32768,
               out: torch.Tensor | None = None):
    """Query expression [N, G] f32 (dense), batch codes [V, N] int32, true types [N]."""
    dev = model.device
    g = _gen(dev, seed + 4)
    comp = torch.randint(0, model.L, (N,), generator=g, device=dev)
    V = len(levels)
    codes = torch.stack([torch.randint(0, lv, (N,), generator=g, device=dev) for lv in levels]) \
        if V else torch.zeros((0, N), dtype=torch.int64, device=dev)
    X = out if out is not None else torch.empty((N, model.G), device=dev, dtype=torch.float32)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        X[s:e] = model.expr(comp[s:e], g, codes[:, s:e] if V else None, seed)
    return X, codes.int(), comp


def type_names(L: int) -> np.ndarray:
    w = max(2, int(math.log10(max(L - 1, 1))) + 1)
    return np.array([f"type_{i:0{w}d}" for i in range(L)], dtype=object)


def level_names(v: int, levels: int) -> np.ndarray:
    # deliberately not zero padded: string order differs from numeric order
    # (`pd.get_dummies` sorts categories as strings)
    return np.array([f"b{v}_{i}" for i in range(levels)], dtype=object)


def to_anndata(model: SynthModel, Zref, ref_comp, C, Nr, X, codes, K: int, R_ref=None,
               missing_frac: float = 0.0, zero_std_frac: float = 0.0, extra_ref_genes: int = 0,
               extra_query_genes: int = 0, sparse: bool = False, seed: int = 0,
               second_label: bool = False, dtype64: bool = True):
    """Wrap generated tensors into a (query, reference) pair of AnnDataLite objects,
    laid out the way `pp.harmony_integrate` leaves a reference (`_utils.py:37-61`).

    missing_frac  : fraction of reference HVGs absent from the query
    zero_std_frac : fraction of reference HVGs with std == 0
    extra_*_genes : non-HVG reference genes / query-only genes, interleaved, so the
                    query needs a real column gather into reference order
    """
    import pandas as pd
    import scipy.sparse as sp

    from .anndata_lite import AnnDataLite

    rng = np.random.default_rng(seed + 11)
    G, d = model.G, model.d
    mean = model.mean.cpu().numpy().astype(np.float64 if dtype64 else np.float32)
    std = model.std.cpu().numpy().astype(np.float64 if dtype64 else np.float32)
    U = model.U.cpu().numpy().astype(np.float64 if dtype64 else np.float32)
    hv_names = np.array([f"g{i}" for i in range(G)], dtype=object)

    # reference var table: HVGs interleaved with non-HVG extras
    n_ref_genes = G + extra_ref_genes
    is_hv = np.zeros(n_ref_genes, dtype=bool)
    hv_pos = np.sort(rng.choice(n_ref_genes, G, replace=False))
    is_hv[hv_pos] = True
    ref_names = np.empty(n_ref_genes, dtype=object)
    ref_names[hv_pos] = hv_names
    ref_names[~is_hv] = [f"refonly{i}" for i in range(extra_ref_genes)]
    r_mean = rng.gamma(2.0, 0.15, n_ref_genes)
    r_std = rng.gamma(2.0, 0.2, n_ref_genes) + 0.05
    r_pcs = rng.normal(size=(n_ref_genes, d))
    r_mean[hv_pos], r_std[hv_pos], r_pcs[hv_pos] = mean, std, U
    if zero_std_frac > 0:
        z = rng.choice(G, max(1, int(zero_std_frac * G)), replace=False)
        r_std[hv_pos[z]] = 0.0
    var = pd.DataFrame({"mean": r_mean.astype(mean.dtype), "std": r_std.astype(std.dtype),
                        "highly_variable": is_hv}, index=pd.Index(ref_names))
    names = type_names(model.L)
    ref_obs = pd.DataFrame({"cell_type": names[ref_comp.cpu().numpy()]},
                           index=pd.Index([f"r{i}" for i in range(Zref.shape[0])]))
    if second_label:
        ref_obs["lineage"] = np.array(["lin_a", "lin_b", "lin_c"], dtype=object)[ref_comp.cpu().numpy() % 3]
    zr = Zref.cpu().numpy().astype(np.float64 if dtype64 else np.float32)
    ref = AnnDataLite(
        sp.csr_matrix((Zref.shape[0], n_ref_genes), dtype=np.float32),
        obs=ref_obs, var=var, obsm={"X_pca": zr, "X_pca_harmony": zr.copy()},
        varm={"PCs": r_pcs.astype(U.dtype)},
        uns={"harmony": {"C": C.cpu().numpy().astype(np.float64), "Nr": Nr.cpu().numpy().astype(np.float64),
                         "K": int(K), "sigma": 0.1, "ref_basis_loadings": "PCs",
                         "ref_basis_adjusted": "X_pca_harmony", "vars_use": None, "harmony_kwargs": {},
                         "converged": True}},
    )
    if R_ref is not None:  # [K, M] as `_utils.py:44-61` stores it
        ref.uns["harmony"]["R"] = R_ref.cpu().numpy().astype(np.float64).T.copy()

    # query var table: a shuffled subset of the HVGs plus query-only genes
    keep = np.ones(G, dtype=bool)
    if missing_frac > 0:
        keep[rng.choice(G, max(1, int(missing_frac * G)), replace=False)] = False
    q_cols = np.flatnonzero(keep)
    n_q_genes = q_cols.size + extra_query_genes
    Xh = X.cpu().numpy()
    if extra_query_genes or missing_frac > 0:
        order = rng.permutation(n_q_genes)
        q_names = np.empty(n_q_genes, dtype=object)
        Xq = np.zeros((Xh.shape[0], n_q_genes), dtype=np.float32)
        q_names[order[: q_cols.size]] = hv_names[q_cols]
        Xq[:, order[: q_cols.size]] = Xh[:, q_cols]
        q_names[order[q_cols.size:]] = [f"qonly{i}" for i in range(extra_query_genes)]
        Xq[:, order[q_cols.size:]] = rng.gamma(2.0, 0.8, (Xh.shape[0], extra_query_genes)) * \
            (rng.random((Xh.shape[0], extra_query_genes)) < 0.08)
    else:
        q_names, Xq = hv_names, Xh
    obs = pd.DataFrame(index=pd.Index([f"q{i}" for i in range(Xh.shape[0])]))
    codes_h = codes.cpu().numpy()
    for v in range(codes_h.shape[0]):
        lv = int(codes_h[v].max()) + 1
        obs[f"batch{v}"] = level_names(v, lv)[codes_h[v]]
    query = AnnDataLite(sp.csr_matrix(Xq) if sparse else Xq, obs=obs,
                        var=pd.DataFrame(index=pd.Index(q_names)), uns={"log1p": {"base": None}})
    return query, ref


def small_case(N=3000, M=10000, G=2000, d=20, K=100, L=12, levels=(1,), seed=0, device="cpu", **kw):
    """BASELINE config 1 by default: tutorial shape, CPU-runnable."""
    model = make_model(G, d, L, seed, device)
    keep_R = kw.pop("keep_R", False)
    parts = make_reference(model, M, K, seed, keep_R=keep_R)
    Zref, comp, C, Nr = parts[:4]
    X, codes, _ = make_query(model, N, list(levels), seed)
    return to_anndata(model, Zref, comp, C, Nr, X, codes, K, R_ref=parts[4] if keep_R else None, seed=seed, **kw)
