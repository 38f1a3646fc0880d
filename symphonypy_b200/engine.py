"""Device-side pipeline: torch tensors in HBM -> C-ABI kernels -> torch tensors in HBM.

PyTorch is used for device memory, streams and (in `_dist.py`) the NCCL
all-reduce; every arithmetic step of the hot path is a call into
`libsymphony_b200.so`.  All functions take CUDA tensors and run on the current
torch stream of the tensor's device.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

CLIP = 10.0  # `symphonypy/_utils.py:253`: max_value, fixed by the reference API
KNN_AUTO, KNN_FFMA, KNN_TCGEN05, KNN_TC16 = 0, 1, 2, 3
KNN_TC = KNN_TCGEN05
KNN_PRUNE_FLAG = 0x100   # SYM_KNN_PRUNE: exact tile pruning, ORed into the method of a tcgen05 scan
PROJECT_TC = os.environ.get("SYM_PROJECT_TC", "1") != "0"   # tensor-core projection when the layout allows


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
    if t.dim() == 2 and t.shape[0] <= 1 and t.stride(0) != t.shape[1]:
        # 0- and 1-row tensors are "contiguous" whatever their row stride is (an empty array from numpy has
        # stride 0); the C ABI takes the row stride as the leading dimension and requires ld >= d
        t = torch.empty(t.shape, device=t.device, dtype=t.dtype).copy_(t)
    return t


def require_cuda(device=None) -> torch.device:
    """The device to run on: `device` if given, else the current CUDA device — under `torchrun` (a process group is
    initialised and LOCAL_RANK is set) the device of this rank, unless the caller already chose one with
    `torch.cuda.set_device`."""
    if not torch.cuda.is_available():
        raise _lib.SymphonyLibraryError(
            "symphonypy_b200 needs a CUDA device (B200, sm_100a); there is no CPU path.")
    if device is not None:
        return torch.device(device)
    cur = torch.cuda.current_device()
    local = os.environ.get("LOCAL_RANK")
    if (cur == 0 and local is not None and torch.distributed.is_available() and torch.distributed.is_initialized()
            and int(local) < torch.cuda.device_count()):
        cur = int(local)
    return torch.device("cuda", cur)


# ------------------------------------------------------------------------------------- project
@dataclass
class Loadings:
    """Replicated reference-side operands of the projection (device resident)."""

    mu: torch.Tensor       # [G] f32
    inv_sd: torch.Tensor   # [G] f32, 0 for genes that contribute nothing
    U: torch.Tensor        # [G, ldu] f32, zero padded, ldu % 4 == 0
    d: int
    z0: torch.Tensor | None = None  # CSR bias, filled lazily
    upack: torch.Tensor | None = None  # loadings packed for the tensor-core projection, filled lazily
    u_scale: float = 1.0               # power of two applied to U in `upack`

    @property
    def G(self) -> int:
        return self.mu.shape[0]


def make_loadings(mean, std, U, present, device) -> Loadings:
    """mean/std/U: host arrays over the G reference genes in use; present: bool [G]
    (`_utils.py:276`: in the query and std != 0)."""
    mean = np.asarray(mean, dtype=np.float64)
    std = np.asarray(std, dtype=np.float64)
    ok = np.asarray(present, dtype=bool) & (std != 0)
    inv = np.zeros_like(std)
    inv[ok] = 1.0 / std[ok]
    U = np.asarray(U, dtype=np.float32)
    G, d = U.shape
    ldu = (d + 3) // 4 * 4
    Up = np.zeros((G, ldu), dtype=np.float32)
    Up[:, :d] = U
    return Loadings(mu=torch.from_numpy(np.where(ok, mean, 0.0).astype(np.float32)).to(device),
                    inv_sd=torch.from_numpy(inv.astype(np.float32)).to(device),
                    U=torch.from_numpy(Up).to(device), d=d)


def project_dense(X: torch.Tensor, ld: Loadings, gene_col: torch.Tensor | None = None,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """Z = clip((X[:, col] - mu) * inv_sd, +-10) @ U   (`_utils.py:298-308`)."""
    lib = _lib.load()
    assert X.is_cuda and X.dim() == 2 and X.dtype == torch.float32 and X.stride(1) == 1
    N = X.shape[0]
    if gene_col is None:
        assert X.shape[1] >= ld.G
    Z = out if out is not None else torch.empty((N, ld.d), device=X.device, dtype=torch.float32)
    # genes already in reference order and 16-byte aligned rows: tensor-core kernel (FP16x3, X-streaming bound);
    # otherwise (column gather, odd strides) the FP32 FFMA kernel
    if PROJECT_TC and gene_col is None and X.stride(0) % 4 == 0 and X.data_ptr() % 16 == 0 and N > 0:
        if ld.upack is None:
            umax = float(ld.U.abs().max().item())
            # power of two bringing max |U| into [512, 1024): the fp16 lo parts stay normal, nothing overflows
            ld.u_scale = float(2.0 ** (9 - math.floor(math.log2(umax)))) if umax > 0 and math.isfinite(umax) else 1.0
            nbytes = int(lib.sym_project_pack_bytes(ld.G, ld.d))
            ld.upack = torch.empty(nbytes, device=X.device, dtype=torch.uint8)
            _lib.check(lib.sym_project_pack(ld.U.data_ptr(), ld.U.stride(0), ld.G, ld.d, ld.u_scale, ld.upack.data_ptr(),
                                            _stream(X.device)), "sym_project_pack")
        _lib.check(lib.sym_project_dense_tc(X.data_ptr(), N, X.stride(0), ld.G, ld.mu.data_ptr(), ld.inv_sd.data_ptr(),
                                            ld.upack.data_ptr(), ld.d, CLIP, 1.0 / ld.u_scale, Z.data_ptr(), Z.stride(0),
                                            _stream(X.device)), "sym_project_dense_tc")
        return Z
    _lib.check(lib.sym_project_dense(X.data_ptr(), N, X.stride(0), _ptr(gene_col), ld.G, ld.mu.data_ptr(),
                                     ld.inv_sd.data_ptr(), ld.U.data_ptr(), ld.U.stride(0), ld.d, CLIP,
                                     Z.data_ptr(), Z.stride(0), _stream(X.device)), "sym_project_dense")
    return Z


def project_csr(data: torch.Tensor, indices: torch.Tensor, indptr: torch.Tensor, N: int, col_gene: torch.Tensor,
                ld: Loadings, out: torch.Tensor | None = None, not_canonical: torch.Tensor | None = None) -> torch.Tensor:
    """CSR form of the same projection; the dense matrix of `_utils.py:240-243` is never built.
    `indptr` may be a slice [s, e] of the matrix's row pointer with `data` / `indices` the WHOLE arrays (its values
    are absolute offsets): rows s..e-1 are then projected into `out` (N = e - s).  `not_canonical` (int32[1], device):
    set to 1 by the kernel when a row's column indices are not strictly increasing."""
    lib = _lib.load()
    dev = data.device
    assert data.dtype == torch.float32 and indices.dtype == torch.int32 and indptr.dtype == torch.int64
    assert indptr.numel() == N + 1
    if ld.z0 is None:
        ld.z0 = torch.empty(ld.d, device=dev, dtype=torch.float32)
        _lib.check(lib.sym_project_csr_bias(ld.mu.data_ptr(), ld.inv_sd.data_ptr(), ld.U.data_ptr(), ld.U.stride(0),
                                            ld.G, ld.d, CLIP, ld.z0.data_ptr(), _stream(dev)), "sym_project_csr_bias")
    Z = out if out is not None else torch.empty((N, ld.d), device=dev, dtype=torch.float32)
    _lib.check(lib.sym_project_csr_checked(data.data_ptr(), indices.data_ptr(), indptr.data_ptr(), N, col_gene.data_ptr(),
                                           ld.mu.data_ptr(), ld.inv_sd.data_ptr(), ld.U.data_ptr(), ld.U.stride(0), ld.d,
                                           CLIP, ld.z0.data_ptr(), Z.data_ptr(), ld.d if N <= 1 else Z.stride(0),
                                           _ptr(not_canonical), _stream(dev)), "sym_project_csr")
    return Z


def assign(Z: torch.Tensor, Y: torch.Tensor, inv_sigma: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """R [N, K] = softmax_k(-2 (1 - cos(z, Y_k)) / sigma_k)   (`_utils.py:154-178`)."""
    lib = _lib.load()
    N, d = Z.shape
    K = Y.shape[0]
    assert Y.shape[1] == d and Y.is_contiguous() and inv_sigma.numel() == K and Z.stride(1) == 1
    R = out if out is not None else torch.empty((N, K), device=Z.device, dtype=torch.float32)
    _lib.check(lib.sym_assign(Z.data_ptr(), Z.stride(0), N, d, Y.data_ptr(), K, inv_sigma.data_ptr(), R.data_ptr(),
                              _stream(Z.device)), "sym_assign")
    return R


# ------------------------------------------------------------------------------------- MoE
@dataclass
class BatchOrder:
    perm: torch.Tensor | None
    seg: torch.Tensor | None
    chunk_start: torch.Tensor | None


def batch_order(codes_v: torch.Tensor, n_levels: int) -> BatchOrder:
    """Counting sort of the cells by level of one batch variable (None when it has one level)."""
    if n_levels == 1:
        return BatchOrder(None, None, None)
    lib = _lib.load()
    dev = codes_v.device
    N = codes_v.shape[0]
    perm = torch.empty(N, device=dev, dtype=torch.int32)
    seg = torch.empty(n_levels + 1, device=dev, dtype=torch.int64)
    chunk_start = torch.empty(n_levels + 1, device=dev, dtype=torch.int32)
    ws_bytes = int(lib.sym_batch_order_workspace(n_levels))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    _lib.check(lib.sym_batch_order(codes_v.data_ptr(), N, n_levels, perm.data_ptr(), seg.data_ptr(),
                                   chunk_start.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev)), "sym_batch_order")
    return BatchOrder(perm, seg, chunk_start)


def moe_stats(Z, R, codes, levels, orders):
    """Sufficient statistics gram [K,B1,B1], cross [K,B1,d] (float64) of `_utils.py:196-204`,
    WITHOUT the priors Nr / C and without the intercept row (both added by `moe_solve`)."""
    lib = _lib.load()
    dev = Z.device
    N, d = Z.shape
    K = R.shape[1]
    V = len(levels)
    B1 = 1 + int(sum(levels))
    offs = np.concatenate([[0], np.cumsum(levels)[:-1]]).astype(np.int32)
    gram = torch.zeros((K, B1, B1), device=dev, dtype=torch.float64)
    cross = torch.zeros((K, B1, d), device=dev, dtype=torch.float64)
    assert codes.dtype == torch.int32 and codes.shape == (V, N) and codes.is_contiguous()
    for v in range(V):
        o = orders[v]
        _lib.check(lib.sym_moe_stats(Z.data_ptr(), Z.stride(0), R.data_ptr(), N, d, K, codes.data_ptr(), V,
                                     offs.ctypes.data, v, int(levels[v]), _ptr(o.perm), _ptr(o.seg),
                                     _ptr(o.chunk_start), B1, gram.data_ptr(), cross.data_ptr(), _stream(dev)),
                   "sym_moe_stats")
    return gram, cross


def moe_solve(gram, cross, Nr, C, lamb_diag, n_levels_var0):
    """W [K,B1,d] f32 = inv(gram + prior + diag(lamb)) @ (cross + prior), row 0 zeroed (`_utils.py:201-209`)."""
    lib = _lib.load()
    dev = gram.device
    K, B1, _ = gram.shape
    d = cross.shape[2]
    W = torch.empty((K, B1, d), device=dev, dtype=torch.float32)
    info = torch.empty(K, device=dev, dtype=torch.int32)
    _lib.check(lib.sym_moe_solve(gram.data_ptr(), cross.data_ptr(), Nr.data_ptr(), C.data_ptr(), lamb_diag.data_ptr(),
                                 K, B1, d, int(n_levels_var0), W.data_ptr(), info.data_ptr(), _stream(dev)),
               "sym_moe_solve")
    return W, info


def moe_apply(Z, R, codes, levels, orders, W, out=None):
    """Zcorr = Z - sum_k R_k W_k^T Phi   (`_utils.py:212`), one pass per batch variable."""
    lib = _lib.load()
    dev = Z.device
    N, d = Z.shape
    K = R.shape[1]
    B1 = W.shape[1]
    Zc = out if out is not None else torch.empty_like(Z)
    src = Z
    off = 0
    for v, lv in enumerate(levels):
        o = orders[v]
        _lib.check(lib.sym_moe_apply(src.data_ptr(), src.stride(0), R.data_ptr(), N, d, K, codes[v].data_ptr(), off,
                                     int(lv), _ptr(o.perm), _ptr(o.seg), _ptr(o.chunk_start), W.data_ptr(), B1,
                                     Zc.data_ptr(), Zc.stride(0), _stream(dev)), "sym_moe_apply")
        src = Zc
        off += int(lv)
    return Zc


# ------------------------------------------------------------------------------------- kNN
@dataclass
class KnnIndex:
    """`KNeighborsClassifier.fit` result: the reference embedding, packed for the scan kernels.

    The packed copy is stored in a coarse spatial order (rows sorted by nearest of C k-means centroids, the
    centroids themselves chained by proximity).  Queries are sorted by the same key at search time and every
    query tile starts its scan at the reference tiles of its own cluster, so its thresholds are tight after a
    few tiles.  The order changes the speed of the scan only, never its result."""

    ref: torch.Tensor      # [M, d] f32, original row order
    packed: torch.Tensor   # opaque bytes
    M: int
    d: int
    centroids: torch.Tensor | None = None   # [C, d] f32, in chain order
    cl_tile: torch.Tensor | None = None     # [C] i32: first 128-row tile of the packed reference with rows of cluster c
    workspace: dict = field(default_factory=dict)
    first_tier: dict = field(default_factory=dict)   # k -> scan method to start with (see knn_search_certified)
    scan_stats: torch.Tensor | None = None           # device [2] int64 of the last pruned scan: tiles scanned, tiles of a full scan


LOCALITY_MIN_REF = 32768     # below this the whole reference is a few hundred tiles: no ordering
LOCALITY_MIN_QUERY = 8192


def nearest_centroid(X: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """code [N] i32 = nearest of the centroids [C, d] (our kernel; used at fit and at search time)."""
    lib = _lib.load()
    N, d = X.shape
    code = torch.empty(N, device=X.device, dtype=torch.int32)
    _lib.check(lib.sym_nearest_centroid(X.data_ptr(), X.stride(0), N, d, centroids.data_ptr(), centroids.shape[0],
                                        code.data_ptr(), _stream(X.device)), "sym_nearest_centroid")
    return code


def _coarse_centroids(ref: torch.Tensor, C: int, iters: int = 4) -> torch.Tensor:
    """A few Lloyd iterations on a sample of the reference (fit time only; the quality of the clustering
    affects scan speed, not results), then the centroids are chained greedily by proximity so that
    neighbouring cluster ids are neighbouring regions."""
    M, d = ref.shape
    g = torch.Generator(device="cpu")
    g.manual_seed(12345)
    sample = ref[torch.randperm(M, generator=g)[: min(M, 128 * C)].to(ref.device)].contiguous()
    if os.environ.get("SYM_KNN_COARSE_INIT", "kpp") == "kpp":
        # k-means++ seeding (sym_kmeanspp, fixed draws): with random starting points Lloyd often leaves two well
        # separated groups of cells in one coarse cluster, whose tiles then mix both — large bounding balls, late
        # thresholds
        gd = torch.Generator(device=ref.device)
        gd.manual_seed(12345)
        first, rand = kmeanspp_draws(sample.shape[0], C, gd, ref.device)
        cent = sample[kmeanspp_ids(sample, C, first, rand).long()].clone()
    else:
        cent = sample[:C].clone()
    for _ in range(iters):
        code = nearest_centroid(sample, cent).long()
        sums = torch.zeros_like(cent).index_add_(0, code, sample)
        cnt = torch.zeros(C, device=ref.device).index_add_(0, code, torch.ones_like(code, dtype=torch.float32))
        cent = torch.where(cnt[:, None] > 0, sums / cnt.clamp_min(1.0)[:, None], cent)
    c = cent.double().cpu().numpy()
    left = np.ones(C, dtype=bool)
    cur = int(np.argmin(c[:, 0]))
    chain = [cur]
    left[cur] = False
    for _ in range(C - 1):
        dist = ((c - c[cur]) ** 2).sum(1)
        dist[~left] = np.inf
        cur = int(np.argmin(dist))
        chain.append(cur)
        left[cur] = False
    return cent[torch.as_tensor(chain, device=ref.device)].contiguous()


def knn_fit(ref: torch.Tensor, locality: bool = True) -> KnnIndex:
    lib = _lib.load()
    ref = _f32c(ref)
    M, d = ref.shape
    centroids = cl_tile = order = None
    if locality and M >= LOCALITY_MIN_REF:
        C = int(min(1024, max(16, M // int(os.environ.get("SYM_KNN_CLUSTER_ROWS", "2048")))))
        centroids = _coarse_centroids(ref, C)
        code = nearest_centroid(ref, centroids)
        bo = batch_order(code, C)          # counting sort of the reference rows by cluster
        order = bo.perm
        cl_tile = (bo.seg[:-1] // 128).to(torch.int32).contiguous()
    nbytes = int(lib.sym_knn_packed_bytes(M, d))
    packed = torch.empty(nbytes, device=ref.device, dtype=torch.uint8)
    assert packed.data_ptr() % 256 == 0
    _lib.check(lib.sym_knn_fit(ref.data_ptr(), ref.stride(0), M, d, _ptr(order), packed.data_ptr(), _stream(ref.device)),
               "sym_knn_fit")
    return KnnIndex(ref=ref, packed=packed, M=M, d=d, centroids=centroids, cl_tile=cl_tile)


def pad_runs(perm: torch.Tensor, seg: torch.Tensor, multiple: int) -> torch.Tensor:
    """`perm` lists row numbers grouped into runs (`seg[c] .. seg[c+1]` = run c, as `batch_order` returns them).
    Returns the list with every non-empty run padded to a multiple of `multiple` entries by repeating its last row."""
    cnt = seg[1:] - seg[:-1]
    pad_cnt = (cnt + multiple - 1) // multiple * multiple
    pad_end = torch.cumsum(pad_cnt, 0)
    pad_off = pad_end - pad_cnt
    n_out = int(pad_end[-1].item()) if pad_end.numel() else 0
    pos = torch.arange(n_out, device=perm.device, dtype=torch.int64)
    run = torch.searchsorted(pad_end, pos, right=True)              # run of each padded position
    within = torch.minimum(pos - pad_off[run], (cnt[run] - 1).clamp_min(0))
    return perm[seg[:-1][run] + within].contiguous()


def knn_search(index: KnnIndex, Q: torch.Tensor, k: int, method: int = KNN_AUTO, want_unsafe: bool = True,
               events: list | None = None, prune: bool = False):
    """idx [N,k] int32 (ascending by (distance, index)), dist2 [N,k] f32, unsafe [N] uint8 | None.
    `events`, if given, receives a (start, end) pair of CUDA events bracketing the scan + re-rank launches.
    `prune` (tcgen05 scans only): skip reference tiles that provably hold no candidate of a whole query CTA — same
    results, data-dependent speed; `index.scan_stats` then holds {tiles scanned, tiles of a full scan}."""
    lib = _lib.load()
    Q = _f32c(Q)
    N, d = Q.shape
    assert d == index.d, f"query has {d} dims, reference has {index.d}"
    dev = Q.device
    idx = torch.empty((N, k), device=dev, dtype=torch.int32)
    dist2 = torch.empty((N, k), device=dev, dtype=torch.float32)
    unsafe = torch.empty(N, device=dev, dtype=torch.uint8) if want_unsafe else None
    q_perm = q_code = cl_tile = None
    if index.centroids is not None and (N >= LOCALITY_MIN_QUERY or prune) and method != KNN_FFMA and N > 0:
        q_code = nearest_centroid(Q, index.centroids)
        bo = batch_order(q_code, index.centroids.shape[0])
        q_perm = bo.perm
        cl_tile = index.cl_tile
        if prune and method in (KNN_TC, KNN_TC16):
            # Pruning bounds a group of 32 consecutive scan rows (one selection warp) by a ball: a group that straddles
            # two clusters of the sorted queries has a ball as wide as the gap between them and holds back its whole CTA.
            # Every cluster's run of rows is therefore padded to a multiple of 32 by repeating its last row (a repeated
            # row is mapped twice and written twice with the same result).
            q_perm = pad_runs(q_perm, bo.seg, 32)
    n_rows = N if q_perm is None else int(q_perm.numel())   # scan rows (more than N only with the padding above)
    ws_bytes = int(lib.sym_knn_workspace(n_rows, index.M, d, k, method))
    key = "ws"
    ws = index.workspace.get(key)
    if ws is None or ws.numel() < ws_bytes or ws.device != dev:
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        index.workspace[key] = ws
    if events is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    flag = KNN_PRUNE_FLAG if (prune and method in (KNN_TC, KNN_TC16) and N > 0) else 0
    _lib.check(lib.sym_knn(Q.data_ptr(), Q.stride(0), n_rows, index.ref.data_ptr(), index.ref.stride(0), index.M,
                           index.packed.data_ptr(), d, k, method | flag, _ptr(q_perm), _ptr(q_code), _ptr(cl_tile),
                           idx.data_ptr(), dist2.data_ptr(), _ptr(unsafe), ws.data_ptr(), ws.numel(), _stream(dev)),
               "sym_knn")
    if events is not None:
        e1.record()
        events.append((e0, e1))
    if flag:
        index.scan_stats = ws[:16].view(torch.int64).clone()
    return idx, dist2, unsafe


EXACT_ROWS_MAX = 64      # = sym_knn_exact_rows_max()


def knn_exact_rows(index: KnnIndex, Q: torch.Tensor, rows: torch.Tensor, k: int, ub: torch.Tensor):
    """Exact float64 brute force for the few query rows `rows` (int64 indices into Q) given upper bounds `ub` [n] f32 on
    their k-th squared distances: (idx [n,k] int32, dist2 [n,k] f32, overflow [n] uint8 — rows the bound did not help)."""
    lib = _lib.load()
    Q = _f32c(Q)
    n = int(rows.numel())
    dev = Q.device
    idx = torch.empty((n, k), device=dev, dtype=torch.int32)
    dist2 = torch.empty((n, k), device=dev, dtype=torch.float32)
    over = torch.empty(n, device=dev, dtype=torch.uint8)
    ws_bytes = int(lib.sym_knn_exact_rows_workspace(n))
    ws = index.workspace.get("exact")
    if ws is None or ws.numel() < ws_bytes or ws.device != dev:
        ws = torch.empty(max(ws_bytes, int(lib.sym_knn_exact_rows_workspace(8))), device=dev, dtype=torch.uint8)
        index.workspace["exact"] = ws
    rows = rows.to(torch.int64).contiguous()
    ub = ub.to(torch.float32).contiguous()
    _lib.check(lib.sym_knn_exact_rows(Q.data_ptr(), Q.stride(0), rows.data_ptr(), n, index.ref.data_ptr(),
                                      index.ref.stride(0), index.M, index.d, k, ub.data_ptr(), idx.data_ptr(),
                                      dist2.data_ptr(), over.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)),
               "sym_knn_exact_rows")
    return idx, dist2, over


def knn_search_certified(index: KnnIndex, Q: torch.Tensor, k: int, method: int = KNN_AUTO, events: list | None = None,
                         prune: bool = False, after_launch=None):
    """`knn_search` plus the repair passes.  Rows whose over-selection margin could not be certified (possible
    with the reduced-precision tensor-core scans, or with many exactly tied neighbours) are re-run with the next
    more accurate scan: single-FP16 tcgen05 -> BF16x3 tcgen05 -> FP32 FFMA.  When only a handful of rows (<= 64) are
    left after the first tier they are finished by exact float64 brute force instead (`knn_exact_rows`, bounded by
    the k-th distance the first tier found for them): a reading of the reference instead of two small-grid launches.

    With `method=KNN_AUTO` the first tier is the single-FP16 scan (a third of the MMAs) unless this index has
    shown, for this k, that more than 8 % of the rows then need the BF16x3 scan anyway (dense neighbourhoods:
    gaps between the k-th and (k+8)-th neighbour below the FP16 score error) — the index remembers and starts
    with BF16x3 from then on.  Results do not depend on the tier.  Returns (idx, dist2, rows re-run)."""
    lib = _lib.load()
    N = Q.shape[0]

    def supported(m: int, n: int) -> bool:
        return bool(lib.sym_knn_method_supported(n, index.M, index.d, k, m))

    adaptive = method == KNN_AUTO
    if adaptive:
        first = index.first_tier.get(k)
        if first is None:
            fp16_on = os.environ.get("SYM_KNN_FP16", "1") != "0"   # (switch for timing experiments)
            first = KNN_TC16 if fp16_on and supported(KNN_TC16, N) else (KNN_TC if supported(KNN_TC, N) else KNN_FFMA)
    else:
        first = method
    idx, dist2, unsafe = knn_search(index, Q, k, first, events=events, prune=prune)
    if after_launch is not None:
        after_launch()   # host work of the caller that can run while the first-tier scan is on the GPU
    n_first = n_bad = int(unsafe.sum().item())
    if n_bad and os.environ.get("SYM_KNN_NO_REPAIR") != "1":  # (the switch exists for kernel timing experiments only)
        rows = torch.nonzero(unsafe, as_tuple=False).flatten()
        if n_bad <= EXACT_ROWS_MAX and os.environ.get("SYM_KNN_EXACT_ROWS", "1") != "0":
            # a handful of rows: exact brute force within the k-th distance the first tier found for them (an upper
            # bound on the true one: uncertified results are still exact distances of real reference rows)
            ub = torch.nextafter(dist2[rows, k - 1] * (1.0 + 1e-6), torch.full((), float("inf"), device=dist2.device))
            i2, d2, over = knn_exact_rows(index, Q, rows, k, ub)
            done = over == 0
            idx[rows[done]] = i2[done]
            dist2[rows[done]] = d2[done]
            rows = rows[~done]                      # (bound useless, e.g. fewer than k survivors: scan tiers below)
            n_bad = int(rows.numel())
        tier = first
        while n_bad and tier != KNN_FFMA:
            tier = KNN_TC if tier == KNN_TC16 and supported(KNN_TC, n_bad) else KNN_FFMA
            i2, d2, u2 = knn_search(index, Q[rows].contiguous(), k, tier)
            idx[rows] = i2
            dist2[rows] = d2
            if tier == KNN_FFMA:
                break
            rows = rows[torch.nonzero(u2, as_tuple=False).flatten()]
            n_bad = int(rows.numel())
    if adaptive and first == KNN_TC16 and N >= 4096 and n_first > 0.08 * N:
        index.first_tier[k] = KNN_TC if supported(KNN_TC, N) else KNN_FFMA
    return idx, dist2, n_first

def vote(idx: torch.Tensor, ref_codes: torch.Tensor, want_conf: bool = True, dist2: torch.Tensor | None = None):
    """Majority vote with sklearn's tie rule (lowest class code); conf = winning fraction.
    With `dist2` (squared distances of the same neighbours) the vote is weighted by 1/distance."""
    lib = _lib.load()
    N, k = idx.shape
    dev = idx.device
    label = torch.empty(N, device=dev, dtype=torch.int32)
    conf = torch.empty(N, device=dev, dtype=torch.float32) if want_conf else None
    assert ref_codes.dtype == torch.int32 and idx.is_contiguous()
    _lib.check(lib.sym_vote(idx.data_ptr(), _ptr(dist2), N, k, ref_codes.data_ptr(), ref_codes.numel(), label.data_ptr(),
                            _ptr(conf), _stream(dev)),
               "sym_vote")
    return label, conf


# ------------------------------------------------------------------------------------- per-cell confidence
def cluster_covariances(X_ref: torch.Tensor, R_ref: torch.Tensor):
    """`_cluster_covs` (`symphonypy/_utils.py:468-496`) on the device: inverse R-weighted covariance
    [K, d, d] and weighted mean [K, d] of every reference cluster, both float64.
    X_ref [M, d] f32, R_ref [K, M] f32 (the layout of `uns["harmony"]["R"]`)."""
    lib = _lib.load()
    X_ref, R_ref = _f32c(X_ref), _f32c(R_ref)
    M, d = X_ref.shape
    K = R_ref.shape[0]
    assert R_ref.shape[1] == M
    dev = X_ref.device
    sums = torch.empty((K, d + 2), device=dev, dtype=torch.float64)
    _lib.check(lib.sym_cluster_moments(X_ref.data_ptr(), X_ref.stride(0), M, d, R_ref.data_ptr(), K, sums.data_ptr(),
                                       _stream(dev)), "sym_cluster_moments")
    v1, v2 = sums[:, 0], sums[:, 1]
    mean = sums[:, 2:] / v1[:, None]
    cov = torch.empty((K, d, d), device=dev, dtype=torch.float64)
    mean32 = mean.float().contiguous()
    _lib.check(lib.sym_cluster_cov(X_ref.data_ptr(), X_ref.stride(0), M, d, R_ref.data_ptr(), K, mean32.data_ptr(),
                                   cov.data_ptr(), _stream(dev)), "sym_cluster_cov")
    cov = cov * (v1 / (v1 * v1 - v2))[:, None, None]          # `_utils.py:493`
    return torch.linalg.inv(cov), mean                          # K tiny d x d inverses (`_utils.py:494`)


def _group_sums(X: torch.Tensor, order: "BatchOrder", n_groups: int):
    """size [G] int64, column sums [G, d] float64 of the rows of X per group of `order`."""
    lib = _lib.load()
    N, d = X.shape
    dev = X.device
    if order.perm is None:
        size = torch.tensor([N], device=dev, dtype=torch.int64)
        ptrs = (None, None, None)
    else:
        size = order.seg[1:] - order.seg[:-1]
        ptrs = (order.perm.data_ptr(), order.seg.data_ptr(), order.chunk_start.data_ptr())
    sums = torch.empty((n_groups, d), device=dev, dtype=torch.float64)
    _lib.check(lib.sym_group_sums(X.data_ptr(), X.stride(0), N, d, n_groups, *ptrs, sums.data_ptr(), _stream(dev)),
               "sym_group_sums")
    return size, sums, ptrs


def group_moments(X: torch.Tensor, codes: torch.Tensor, n_groups: int):
    """Per-group size [G] (int64), mean [G, d] and centred scatter matrix sum (x-mean)(x-mean)^T [G, d, d]
    (float64) of the rows of X grouped by `codes` in [0, G) — the per-slice numpy of
    `symphonypy/_utils.py:436-444`, for all groups in two passes over X."""
    lib = _lib.load()
    X = _f32c(X)
    N, d = X.shape
    dev = X.device
    assert codes.dtype == torch.int32 and codes.shape == (N,)
    order = batch_order(codes, n_groups)
    size, sums, ptrs = _group_sums(X, order, n_groups)
    mean = sums / size.clamp_min(1)[:, None].double()
    mean32 = mean.float().contiguous()
    scatter = torch.empty((n_groups, d, d), device=dev, dtype=torch.float64)
    _lib.check(lib.sym_group_cov(X.data_ptr(), X.stride(0), N, d, n_groups, *ptrs, mean32.data_ptr(),
                                 scatter.data_ptr(), _stream(dev)), "sym_group_cov")
    return size, mean, scatter


# ------------------------------------------------------------------------------------- k-means (no-Harmony fallback)
def kmeanspp_draws(N: int, K: int, gen: torch.Generator, dev):
    """The random numbers one greedy k-means++ seeding consumes: the first centre's row and, for each further centre,
    2 + int(ln K) uniform draws (sklearn's `n_local_trials`)."""
    trials = 2 + int(np.log(K))
    first = int(torch.randint(N, (1,), generator=gen, device=dev).item())
    rand = torch.rand((max(K - 1, 1), trials), generator=gen, device=dev, dtype=torch.float64)
    return first, rand


def kmeanspp_ids(X: torch.Tensor, K: int, first: int, rand: torch.Tensor) -> torch.Tensor:
    """Rows of X chosen by greedy k-means++ (`sym_kmeanspp`) for the given draws: [K] int32."""
    lib = _lib.load()
    X = _f32c(X)
    N, d = X.shape
    dev = X.device
    trials = int(rand.shape[1])
    assert rand.dtype == torch.float64 and rand.is_contiguous() and rand.shape[0] >= K - 1
    ws = torch.empty(int(lib.sym_kmeanspp_workspace(N, trials)), device=dev, dtype=torch.uint8)
    ids = torch.empty(K, device=dev, dtype=torch.int32)
    _lib.check(lib.sym_kmeanspp(X.data_ptr(), X.stride(0), N, d, K, trials, int(first), rand.data_ptr(), ids.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream(dev)), "sym_kmeanspp")
    return ids


def _kmeanspp(X: torch.Tensor, K: int, gen: torch.Generator) -> torch.Tensor:
    """Greedy k-means++ centres [K, d] f32 (sklearn's `_kmeans_plusplus`), chosen on the device."""
    first, rand = kmeanspp_draws(X.shape[0], K, gen, X.device)
    return X[kmeanspp_ids(X, K, first, rand).long()].contiguous()


def kmeans(X: torch.Tensor, K: int, n_init: int = 10, max_iter: int = 25, tol: float = 1e-4, seed: int | None = None):
    """Lloyd k-means, k-means++ seeding, best of `n_init` runs by inertia: the estimator
    `KMeans(n_clusters=K, init="k-means++", n_init=10, max_iter=25)` of `symphonypy/_utils.py:323-325`.
    Assignment step = `sym_nearest_centroid`, update step = `sym_batch_order` + `sym_group_sums` (FP64 sums);
    a run stops when the squared centre shift falls below tol * mean feature variance (sklearn's rule).
    Returns (centres [K, d] float64, inertia)."""
    X = _f32c(X)
    N, d = X.shape
    assert 0 < K <= N, "kmeans: need 1 <= K <= number of rows"
    dev = X.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) if seed is not None else int.from_bytes(os.urandom(4), "little"))
    Xd = X.double()
    x2 = (Xd ** 2).sum(1)
    total_x2 = x2.sum()
    tol_abs = float(tol * Xd.var(dim=0, unbiased=False).mean())
    best_inertia, best_cent = None, None
    for _ in range(max(1, n_init)):
        cent32 = _kmeanspp(X, K, gen)
        cent = cent32.double()
        inertia = None
        for _it in range(max_iter):
            code = nearest_centroid(X, cent32)
            size, sums, _ = _group_sums(X, batch_order(code, K), K)
            new = torch.where(size[:, None] > 0, sums / size.clamp_min(1)[:, None].double(), cent)  # empty: keep
            shift = float(((new - cent) ** 2).sum())
            # sum ||x - mean of its cluster||^2 = sum ||x||^2 - sum n_g ||mean_g||^2
            inertia = float(total_x2 - (size.double() * (new ** 2).sum(1) * (size > 0)).sum())
            cent, cent32 = new, new.float().contiguous()
            if shift <= tol_abs:
                break
        if best_inertia is None or inertia < best_inertia:
            best_inertia, best_cent = inertia, cent
    return best_cent, best_inertia

def cell_confidence(Xq: torch.Tensor, Rq: torch.Tensor, mean: torch.Tensor, inv_cov: torch.Tensor) -> torch.Tensor:
    """out[n] = sum_k Rq[n,k] * Mahalanobis(x_n; mean_k, cov_k)   (`symphonypy/tl.py:88-111`)."""
    lib = _lib.load()
    Xq, Rq = _f32c(Xq), _f32c(Rq)
    N, d = Xq.shape
    K = Rq.shape[1]
    mean32, inv32 = _f32c(mean), _f32c(inv_cov)
    out = torch.empty(N, device=Xq.device, dtype=torch.float32)
    _lib.check(lib.sym_cell_confidence(Xq.data_ptr(), Xq.stride(0), N, d, Rq.data_ptr(), K, mean32.data_ptr(),
                                       inv32.data_ptr(), out.data_ptr(), _stream(Xq.device)), "sym_cell_confidence")
    return out
