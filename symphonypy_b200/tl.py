"""Drop-in `tl` namespace: `map_embedding`, `transfer_labels_kNN`, `per_cell_confidence` and
`per_cluster_confidence` on B200.

Mirrors `symphonypy/tl.py:271-436` — same signatures, defaults, `obsm` / `obs` keys,
assertion messages, warnings and logger name — with the numerics of
`symphonypy/_utils.py:154-308` and scikit-learn's brute-force `KNeighborsClassifier`
replaced by the CUDA kernels behind `include/symphony_b200.h`.

Differences a user can observe (all documented in DESIGN.md):
  * arrays written to `obsm` are float32 (the kernels compute in FP32; tolerance 1e-4
    relative against the reference's float64);
  * the no-Harmony fallback (`tl.py:320-335`, `_utils.py:311-352`) clusters the reference with a GPU
    k-means; like the reference's unseeded sklearn call its centres are random;
  * kNN options the kernels do not implement (non-Euclidean metric, callable weights)
    raise `NotImplementedError` — there is no CPU fallback;
  * opt-in extras (keyword only): `confidence_key=` / `neighbors_key=` on
    `transfer_labels_kNN`;
  * under `torch.distributed` (one process per GPU) every rank passes its own shard of
    the query; the MoE statistics are all-reduced so results equal the single-process run.
"""
from __future__ import annotations

import logging
import numbers
import os
import warnings
import weakref
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pandas as pd
import torch
from scipy.sparse import issparse

from . import _design, _dist, _lib, engine, pipeline

logger = logging.getLogger("symphonypy")


class settings:
    """Process-wide knobs of the B200 backend."""

    device = None                  # torch device; default: current CUDA device (LOCAL_RANK under torchrun)
    h2d_chunk_bytes = 256 << 20    # rows of X are streamed to the GPU in chunks of about this size
    knn_method = engine.KNN_AUTO   # 0 auto (tiered), 1 FP32 FFMA scan, 2 tcgen05 BF16x3 scan, 3 tcgen05 single-FP16 scan
    # Remember device copies of the arrays this module wrote to obsm (saves their re-upload in transfer_labels_kNN /
    # per_cell_confidence).  Off by default: a remembered copy is found by the identity of the numpy array, so an
    # IN-PLACE edit of that obsm array between two calls would go unnoticed.  Re-uploading from the pinned result
    # buffer costs ~2 ms per 100 MB.
    keep_on_device = False
    # kNN: skip reference tiles that provably hold no candidate for a whole tile of query cells (bounding balls of the
    # coarse-cluster-ordered reference).  Results are identical (certified exact); the gain depends on how clustered
    # the embedding is (none for unclustered data).  Off by default: the benchmark headline is the full scan.
    knn_prune = False
    # Opt-in: page-lock large host input arrays IN PLACE (cudaHostRegister) the second time the same array is passed
    # in; from then on it is DMA'd straight from where it lies instead of being staged through pinned buffers by the
    # host's threads.  Registration costs ~0.4 s per GB once (measured) and saves a copy at memory speed per call, so it
    # only pays for workflows that map the same arrays many times, or when few host threads are available (torchrun's
    # OMP_NUM_THREADS=1).  Arrays the CALLER allocated in page-locked memory are always used directly.  The registration
    # is dropped when the array is garbage collected or by `clear_caches()`.
    pin_repeated_inputs = False
    store_R = True                 # write obsm[f"{basis}_symphony_R"] (N x K) as the reference does
    pinned_pool_bytes = 2 << 30    # page-locked result buffers kept for reuse after their arrays are collected


# ----------------------------------------------------------------------------- device caches
# Reference-side device objects (loadings, packed kNN index) are cached across calls: building them costs more than
# mapping a query.  Every cache key carries a DIGEST OF THE FULL CONTENT of the host arrays the object was built
# from (mean, std, gene mask, loadings; the reference embedding), so the normal scanpy idiom of editing
# `adata_ref` in place (re-scaling, re-running PCA, replacing rows) is noticed and the object rebuilt.  Hashing
# runs at memory speed on a few threads (120 MB of float64 embedding: ~4 ms) — far below a rebuild.
# Label columns are NOT cached: they are re-encoded on every call (3-20 ms for 500k labels).
_DEV_CACHE: "OrderedDict[int, tuple]" = OrderedDict()   # id(ndarray) -> (weakref, cuda tensor); settings.keep_on_device
_REF_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
# (id(reference embedding array), device) -> (weakref to the array, content digest, kNN index, (shape, dtype, address)):
# the index searched speculatively while the digest of the array is recomputed (see transfer_labels_kNN)
_KNN_SPEC: "OrderedDict[tuple, tuple]" = OrderedDict()

_DIGEST_POOL: list = []


def _hash_threads() -> int:
    """Host threads one digest may use (one process per GPU on one host: the ranks share its cores)."""
    ranks = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1) if _dist.active() else 1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(8, cores // max(1, ranks)))


def content_digest(*arrays) -> str:
    """64-bit digests of the full contents (plus dtype and shape) of host arrays, as one string.  The bytes are
    hashed by `sym_host_digest64` of the C-ABI library on a few threads: the call releases the interpreter lock, so a
    digest computed on a worker thread does not hold up the thread that is launching GPU work."""
    out = []
    for a in arrays:
        if a is None:
            out.append("-")
            continue
        if isinstance(a, torch.Tensor):
            out.append(f"t{a.data_ptr():x}:{tuple(a.shape)}:{a._version}")   # device tensors: identity + version counter
            continue
        a = np.asarray(a)
        if a.dtype == object:
            a = np.asarray(pd.util.hash_pandas_object(pd.Series(a), index=False))
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        h = int(_lib.load().sym_host_digest64(a.ctypes.data if a.size else None, a.nbytes, _hash_threads()))
        out.append(f"{a.dtype.str}{a.shape}{h:x}")
    return "|".join(out)


def _remember(arr: np.ndarray, t: torch.Tensor) -> None:
    if not settings.keep_on_device:
        return
    key = id(arr)
    _DEV_CACHE[key] = (weakref.ref(arr, lambda _r, k=key: _DEV_CACHE.pop(k, None)), t)
    while len(_DEV_CACHE) > 8:
        _DEV_CACHE.popitem(last=False)


def _device_copy(arr, dev) -> torch.Tensor:
    """Device float32 copy of an obsm entry (cache hit only with `settings.keep_on_device`)."""
    if isinstance(arr, torch.Tensor):
        return arr.to(dev, torch.float32)
    if settings.keep_on_device:
        hit = _DEV_CACHE.get(id(arr))
        if hit is not None and hit[0]() is arr and hit[1].device == dev:
            return hit[1]
    a = np.ascontiguousarray(np.asarray(arr))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # (read-only arrays: torch warns that it could write through the view)
        t = torch.from_numpy(a)
    # a page-locked source (the result buffers `map_embedding` hands out are) is copied asynchronously: the array stays
    # referenced by the caller's AnnData, and everything that reads the copy is ordered behind it on the stream
    return t.to(dev, non_blocking=t.is_pinned()).float()


def _cached(key, build):
    """`build()` once per key.  Keys hold content digests, so there is nothing to invalidate; old entries fall out
    of the 8-entry LRU (or `clear_caches()`)."""
    hit = _REF_CACHE.get(key)
    if hit is not None:
        _REF_CACHE.move_to_end(key)
        return hit
    val = build()
    _REF_CACHE[key] = val
    while len(_REF_CACHE) > 8:
        _REF_CACHE.popitem(last=False)
    return val


def clear_caches() -> None:
    """Forget every cached device object, staging buffer and pooled page-locked result buffer."""
    _DEV_CACHE.clear()
    _REF_CACHE.clear()
    _KNN_SPEC.clear()
    _SEEN_INPUTS.clear()
    for ptr in list(_PINNED_IN_PLACE):
        _unregister(ptr)
    _PIN_BUFS.clear()
    _POOL.clear()


# ----------------------------------------------------------------------------- project
_PIN_BUFS: dict = {}


def _pinned_stage(nbytes: int, slot: int) -> torch.Tensor:
    buf = _PIN_BUFS.get(slot)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        _PIN_BUFS[slot] = buf
    return buf


# Host -> device streaming.  A cudaMemcpy out of pageable memory is staged by the driver at a few GB/s on one
# thread; here the next chunk is copied (and, where needed, converted to the 32-bit type the kernels read) into a
# page-locked ring by the host's threads while the previous chunk is on the PCIe bus, so that pageable numpy / scipy
# inputs stream at close to the speed of pinned ones.
_STAGE_CHUNK = 48 << 20      # bytes per staged chunk


_TORCH_NP = {torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64, torch.float64: np.float64,
             torch.uint8: np.uint8}


def _stage_copy(dst: torch.Tensor, src: np.ndarray) -> None:
    """dst[...] = src for host arrays (1-D, same length; converts the dtype on the way).  torch's CPU copy is
    parallel over its intra-op threads and runs at memory speed (a thread pool of numpy copies measured half of it).
    When the process has been left fewer intra-op threads than its share of the cores (torchrun exports
    OMP_NUM_THREADS=1: one staging thread per rank, 4x slower end to end) a same-dtype copy goes through the library's
    own threads instead (`sym_host_copy`)."""
    want = _hash_threads()
    if (want > torch.get_num_threads() and src.nbytes >= (1 << 20) and src.flags.c_contiguous and dst.is_contiguous()
            and _TORCH_NP.get(dst.dtype) == src.dtype.type and dst.numel() == src.size):
        _lib.load().sym_host_copy(dst.data_ptr(), src.ctypes.data, src.nbytes, want)
        return
    if not src.flags.writeable:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dst.copy_(torch.from_numpy(src))
    else:
        dst.copy_(torch.from_numpy(src))


_PINNED_IN_PLACE: dict = {}                       # data pointer -> nbytes (arrays this module page-locked in place)
_SEEN_INPUTS: "OrderedDict[tuple, bool]" = OrderedDict()
_PIN_MIN_BYTES = 8 << 20


def _unregister(ptr: int) -> None:
    if _PINNED_IN_PLACE.pop(ptr, None) is not None:
        try:
            torch.cuda.cudart().cudaHostUnregister(ptr)
        except Exception:   # interpreter shutdown, context gone
            pass


def _host_tensor(arr: np.ndarray) -> torch.Tensor:
    if arr.flags.writeable:
        return torch.from_numpy(arr)
    with warnings.catch_warnings():   # torch warns on read-only memory; it is only read
        warnings.simplefilter("ignore")
        return torch.from_numpy(arr)


def _pinned_in_place(arr: np.ndarray) -> bool:
    """True when `arr` lies in page-locked memory — allocated so by the caller, or registered here because this is
    (at least) the second call that passes the same array."""
    if not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous or arr.nbytes == 0:
        return False
    try:
        if _host_tensor(arr).is_pinned():
            return True
    except Exception:
        return False
    if not settings.pin_repeated_inputs or arr.nbytes < _PIN_MIN_BYTES:
        return False
    ptr = arr.ctypes.data
    key = (ptr, arr.nbytes)
    if key not in _SEEN_INPUTS:
        _SEEN_INPUTS[key] = True
        while len(_SEEN_INPUTS) > 64:
            _SEEN_INPUTS.popitem(last=False)
        return False
    try:
        rc = torch.cuda.cudart().cudaHostRegister(ptr, arr.nbytes, 0)
        if int(rc) != 0:
            return False
        _PINNED_IN_PLACE[ptr] = arr.nbytes
        weakref.finalize(arr, _unregister, ptr)
    except Exception:
        return False
    return True


class _Stager:
    """Ring of page-locked staging buffers + a copy stream.  `send(pairs)` stages each (host 1-D array -> device 1-D
    tensor of a 32-bit dtype) pair of one chunk and enqueues its H2D; returns the event the compute stream waits on."""

    def __init__(self, dev, nbuf: int = 3):
        self.dev = dev
        self.copy = torch.cuda.Stream(dev)
        self.copy.wait_stream(torch.cuda.current_stream(dev))
        self.nbuf = nbuf
        self.done = [None] * nbuf
        self.i = 0

    def send(self, pairs, direct=None) -> torch.cuda.Event:
        b = self.i % self.nbuf
        self.i += 1
        if self.done[b] is not None:
            self.done[b].synchronize()     # the H2D that last read this staging buffer has finished
        need = sum(((dst.numel() * 4 + 255) & ~255) for _, dst in pairs)
        stage = _pinned_stage(max(need, _STAGE_CHUNK + (1 << 20)), ("stager", b))
        off = 0
        with torch.cuda.stream(self.copy):
            for j, (src, dst) in enumerate(pairs):
                n = dst.numel()
                if n == 0:
                    continue
                if direct is not None and direct[j]:   # already page-locked, already the right type: no staging
                    dst.copy_(_host_tensor(src), non_blocking=True)
                    continue
                view = stage[off:off + n * 4].view(dst.dtype)
                _stage_copy(view, src)
                dst.copy_(view, non_blocking=True)
                off += (n * 4 + 255) & ~255
            ev = torch.cuda.Event()
            ev.record(self.copy)
        self.done[b] = ev
        return ev


def _project_host_dense(Xh: torch.Tensor, ld: engine.Loadings, gene_col, Z: torch.Tensor, after_chunk=None,
                        host_array: np.ndarray | None = None) -> None:
    """Stream a host matrix through the GPU: H2D of chunk i+1 overlaps the projection of chunk i.
    `after_chunk(Z, s, e)` runs on the compute stream right after rows [s, e) have been projected.
    Page-locked float32 input goes straight onto the bus; anything else (pageable, float64, integer counts) is
    staged — and converted to float32 on the way — by `_Stager`."""
    dev = Z.device
    N, C = Xh.shape
    if N == 0:
        return
    # (`host_array`: the caller's own ndarray object behind Xh — the in-place registration lives as long as it does)
    direct = Xh.dtype == torch.float32 and (Xh.is_pinned() or (host_array is not None and _pinned_in_place(host_array)))
    row_bytes = C * 4
    chunk = settings.h2d_chunk_bytes if direct else _STAGE_CHUNK
    rows = int(max(1, min(N, chunk // max(1, row_bytes))))
    compute = torch.cuda.current_stream(dev)
    bufs = [torch.empty((rows, C), dtype=torch.float32, device=dev) for _ in range(2 if N > rows else 1)]
    free = [None for _ in bufs]
    if direct:
        copy = torch.cuda.Stream(dev)
        copy.wait_stream(compute)
    else:
        stager = _Stager(dev)
        Xn = Xh.numpy()
    for i, s in enumerate(range(0, N, rows)):
        b = i % len(bufs)
        e = min(N, s + rows)
        if direct:
            with torch.cuda.stream(copy):
                if free[b] is not None:
                    copy.wait_event(free[b])
                bufs[b][: e - s].copy_(Xh[s:e], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy)
        else:
            if free[b] is not None:
                stager.copy.wait_event(free[b])
            ready = stager.send([(Xn[s:e].reshape(-1), bufs[b][: e - s].view(-1))])
        compute.wait_event(ready)
        engine.project_dense(bufs[b][: e - s], ld, gene_col, out=Z[s:e])
        if after_chunk is not None:
            after_chunk(Z, s, e)
        free[b] = torch.cuda.Event()
        free[b].record(compute)


def csr_row_chunks(indptr: np.ndarray, per_chunk: int) -> list:
    """Row boundaries [0, r1, r2, ..., N] of a CSR matrix such that every chunk of rows holds about `per_chunk`
    non-zeros (never more, unless a single row does) and at least one row."""
    N = len(indptr) - 1
    bounds = [0]
    while bounds[-1] < N:
        target = indptr[bounds[-1]] + per_chunk
        nxt = int(np.searchsorted(indptr, target, side="right")) - 1
        bounds.append(min(N, max(nxt, bounds[-1] + 1)))
    return bounds


def _project_host_csr(Xc, col_gene: np.ndarray, ld: engine.Loadings, Z: torch.Tensor, after_chunk=None):
    """Stream a scipy CSR matrix through the GPU in row chunks (its `data` / `indices` slices are contiguous): the
    values and column indices of chunk i+1 are staged and sent while chunk i is projected; `after_chunk(Z, s, e)`
    follows each chunk on the compute stream.  Returns the device flag "some row is not in canonical format"
    (None when scipy already knows that the matrix is canonical)."""
    dev = Z.device
    N = Xc.shape[0]
    indptr_h = np.ascontiguousarray(Xc.indptr)
    nnz = int(indptr_h[-1]) if N > 0 else 0
    compute = torch.cuda.current_stream(dev)
    data = torch.empty(nnz, device=dev, dtype=torch.float32)
    indices = torch.empty(nnz, device=dev, dtype=torch.int32)
    indptr = torch.from_numpy(indptr_h.astype(np.int64, copy=False)).to(dev)
    colg = torch.from_numpy(col_gene).to(dev)
    known = getattr(Xc, "_has_canonical_format", None) is True     # scipy's cached flag; never computed here (O(nnz))
    flag = None if known else torch.zeros(1, device=dev, dtype=torch.int32)
    if N == 0:
        return flag
    stager = _Stager(dev)
    per_chunk = max(1, _STAGE_CHUNK // 8)                          # non-zeros per chunk (4 B value + 4 B index)
    bounds = csr_row_chunks(indptr_h, per_chunk)
    d_h, i_h = Xc.data, Xc.indices
    # values / indices that are page-locked (by the caller, or in place on their second visit) and already 32-bit
    # skip the staging copy
    direct = (d_h.dtype == np.float32 and _pinned_in_place(d_h), i_h.dtype == np.int32 and _pinned_in_place(i_h))
    for s, e in zip(bounds[:-1], bounds[1:]):
        lo, hi = int(indptr_h[s]), int(indptr_h[e])
        for o in range(lo, hi, per_chunk):                         # (a single row wider than a chunk: several sends)
            o2 = min(hi, o + per_chunk)
            ready = stager.send([(d_h[o:o2], data[o:o2]), (i_h[o:o2], indices[o:o2])], direct)
        if hi > lo:
            compute.wait_event(ready)
        engine.project_csr(data, indices, indptr[s:e + 1], e - s, colg, ld, out=Z[s:e], not_canonical=flag)
        if after_chunk is not None:
            after_chunk(Z, s, e)
    return flag


class _NotCanonical(Exception):
    """The CSR query has unsorted or duplicated column indices (found by the projection kernel)."""


def _project(adata_query, adata_ref, loadings_key, use_genes_column, dev, after_chunk=None, X=None,
             pending_checks: list | None = None) -> torch.Tensor:
    """`_map_query_to_ref` (`_utils.py:248-308`): returns Z [N, d] on the device.  `after_chunk(Z, s, e)` is called
    on the compute stream whenever rows [s, e) of Z are final (once per H2D chunk for a host matrix).
    `X` overrides `adata_query.X`; `pending_checks` receives device flags the caller reads at its next sync point
    (CSR input: "not in canonical format" — the caller then canonicalises a copy and calls again)."""
    if pending_checks is None:
        pending_checks = []
    assert "mean" in adata_ref.var, "Gene expression means are expected to be saved in adata_ref.var"
    assert "std" in adata_ref.var, "Gene expression stds are expected to be saved in adata_ref.var"
    if use_genes_column is None:
        hv = None
        genes = adata_ref.var_names
        stds = np.array(adata_ref.var["std"])
    else:
        hv = np.asarray(adata_ref.var[use_genes_column]).astype(bool)
        genes = adata_ref.var_names[hv]
        stds = np.array(adata_ref.var["std"])[hv]
    q_names = adata_query.var_names
    q_idx = q_names.get_indexer(genes)
    present = (q_idx >= 0) & (stds != 0)  # `_utils.py:276`
    if not present.all():  # same message as `_utils.py:233-239`
        logger.warning(
            "%i out of %i genes from the reference are missing in the query dataset or have zero std in the reference, "
            "their expressions in the query will be set to zero", int((~present).sum()), genes.shape[0])

    means_all = np.asarray(adata_ref.var["mean"])
    pcs = np.asarray(adata_ref.varm[loadings_key])

    def build():
        means = np.array(means_all) if hv is None else np.array(means_all)[hv]
        U = pcs if hv is None else pcs[np.asarray(adata_ref.var_names.isin(genes))]  # `_utils.py:307`
        return engine.make_loadings(means, stds, U, present, dev)

    # key = full content of everything the device operands are built from (in-place edits of adata_ref are noticed)
    ld = _cached(("loadings", str(dev), content_digest(means_all, stds, hv, present, pcs)), build)
    G = ld.G
    if X is None:
        X = adata_query.X
    N = X.shape[0]
    Z = torch.empty((N, ld.d), device=dev, dtype=torch.float32)
    identity = bool((q_idx == np.arange(G)).all())

    if issparse(X):
        Xc = X.tocsr()     # (no copy when X already is CSR)
        col_gene = np.full(X.shape[1], -1, dtype=np.int32)
        ok = q_idx >= 0
        col_gene[q_idx[ok]] = np.arange(G, dtype=np.int32)[ok]
        flag = _project_host_csr(Xc, col_gene, ld, Z, after_chunk)
        if flag is not None:
            pending_checks.append(flag)
        return Z

    gene_col = None if identity else torch.from_numpy(q_idx.astype(np.int32)).to(dev)
    Xn = None
    if isinstance(X, torch.Tensor):
        if X.is_cuda:
            xd = X if (X.dtype == torch.float32 and X.stride(1) == 1) else X.float().contiguous()
            engine.project_dense(xd, ld, gene_col, out=Z)
            if after_chunk is not None and N > 0:
                after_chunk(Z, 0, N)
            return Z
        Xh = X if X.is_contiguous() else X.contiguous()
    else:
        Xn = np.asarray(X)
        if not Xn.flags.c_contiguous:
            Xn = np.ascontiguousarray(Xn)
        if not Xn.flags.writeable:  # torch.from_numpy warns on read-only memory; we never write to it
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                Xh = torch.from_numpy(Xn)
        else:
            Xh = torch.from_numpy(Xn)
    _project_host_dense(Xh, ld, gene_col, Z, after_chunk, host_array=Xn if Xn is X else None)
    return Z


# ----------------------------------------------------------------------------- map_embedding
def _inv_sigma(sigma, K: int, dev) -> torch.Tensor:
    """`_utils.py:158-164`: sigma is cast to float32 first (so 0.1 is float32(0.1))."""
    if np.ndim(sigma) == 0:
        s = np.full(K, np.float32(sigma), dtype=np.float32)
    else:
        s = np.array(sigma, dtype=np.float32)
        assert len(s) == K, \
            "sigma paramater must be either a single float or an array of length equal to number of clusters"
    return torch.from_numpy((1.0 / s.astype(np.float64)).astype(np.float32)).to(dev)


class _PinnedPool:
    """Page-locked host buffers for results.  A device->host copy into pageable memory runs at a few GB/s;
    into pinned memory it runs at PCIe speed and asynchronously.  The numpy arrays handed to the user are
    views of pool buffers; a buffer goes back to the pool when its array is garbage collected.

    Buffers are sized to the request rounded up to 2 MiB (not to a power of two: an N x K result of 10M cells
    must not pin twice its size), and the pool keeps at most `settings.pinned_pool_bytes` of FREE buffers —
    what does not fit is released to the OS instead of staying page-locked."""

    GRAIN = 2 << 20

    def __init__(self):
        self.free: dict[int, list] = {}
        self.free_bytes = 0

    def take(self, nbytes: int) -> torch.Tensor:
        size = max(self.GRAIN, (int(nbytes) + self.GRAIN - 1) // self.GRAIN * self.GRAIN)
        bucket = self.free.get(size)
        if bucket:
            self.free_bytes -= size
            return bucket.pop()
        return torch.empty(size, dtype=torch.uint8).pin_memory()

    def give(self, buf: torch.Tensor) -> None:
        size = buf.numel()
        if self.free_bytes + size > settings.pinned_pool_bytes:
            return   # dropped: the page-locked memory is freed with the tensor
        self.free.setdefault(size, []).append(buf)
        self.free_bytes += size

    def clear(self) -> None:
        self.free.clear()
        self.free_bytes = 0


_POOL = _PinnedPool()


def _to_host_async(t: torch.Tensor) -> np.ndarray:
    """Starts a D2H copy of `t` on the current stream into pinned memory and returns the numpy view;
    the caller synchronises the stream before handing the array out."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return np.empty(tuple(t.shape), dtype=np.float32 if t.dtype == torch.float32 else np.int32)
    buf = _POOL.take(nbytes)
    host = buf[:nbytes].view(t.dtype).view(t.shape)
    host.copy_(t, non_blocking=True)
    arr = host.numpy()
    weakref.finalize(arr, _POOL.give, buf)
    return arr


def _pinned_result(shape, dtype=torch.float32):
    """(pinned host tensor, numpy view of it) for a result that is filled by asynchronous D2H copies; the buffer
    returns to the pool when the numpy array is garbage collected.  (None, empty array) for empty results."""
    n = int(np.prod(shape))
    if n == 0:
        return None, np.empty(tuple(shape), dtype=np.float32 if dtype == torch.float32 else np.int32)
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    buf = _POOL.take(nbytes)
    host = buf[:nbytes].view(dtype).view(tuple(shape))
    arr = host.numpy()
    weakref.finalize(arr, _POOL.give, buf)
    return host, arr


def _to_host(t: torch.Tensor) -> np.ndarray:
    arr = _to_host_async(t)
    torch.cuda.current_stream(t.device).synchronize()
    return arr


# ----------------------------------------------------------------------------- no-Harmony fallback
def _run_soft_kmeans(adata_ref, ref_basis: str = "X_pca", K: int | None = None, ref_basis_loadings: str = "PCs",
                     sigma: float = 0.1, random_state: int | None = None) -> None:
    """Reference summary for a reference that was never Harmony-integrated (`symphonypy/_utils.py:311-352`):
    k-means (k-means++, best of 10 runs, 25 iterations) on `obsm[ref_basis]`, soft assignment of the reference
    cells to the unit-norm centres, and the same `uns["harmony"]` fields.  Quirks kept: `C` holds the k-means
    centres (not the R-weighted sums Harmony stores) while `Nr` is the soft count.

    Like the reference (which does not seed sklearn's KMeans) the result is random unless `random_state` is
    given; the clustering runs on the GPU (`engine.kmeans`)."""
    N = adata_ref.shape[0]
    if K is None:
        K = np.min([np.round(N / 30.0), 100]).astype(int)
    K = int(K)
    dev = engine.require_cuda(settings.device)
    with torch.cuda.device(dev):
        Z = _device_copy(adata_ref.obsm[ref_basis], dev)
        # Under torch.distributed every rank holds the same reference but would draw its own random centres: the
        # all-reduced MoE statistics would then mix unrelated clusterings.  Rank 0 clusters, everyone gets its centres.
        if _dist.rank() == 0:
            C, _ = engine.kmeans(Z, K, n_init=10, max_iter=25, seed=random_state)
        else:
            C = torch.empty((K, Z.shape[1]), device=dev, dtype=torch.float64)
        C = _dist.broadcast(C.contiguous(), src=0)
        Y = C / C.norm(dim=1, keepdim=True)
        R = engine.assign(Z, Y.float().contiguous(), _inv_sigma(sigma, K, dev))     # [N, K]
        Rt = R.double().T.contiguous()                                             # stored [K, N], as Harmony's R
        adata_ref.uns["harmony"] = {
            "Nr": Rt.sum(dim=1).cpu().numpy(),
            "C": C.cpu().numpy(),
            "K": K,
            "sigma": None,
            "ref_basis_loadings": ref_basis_loadings,
            "ref_basis_adjusted": ref_basis,
            "vars_use": None,
            "harmony_kwargs": {},
            "converged": True,
            "R": Rt.cpu().numpy(),
        }


def map_embedding(
    adata_query,
    adata_ref,
    key: list[str] | str | None = None,
    lamb: float | np.ndarray | None = None,
    sigma: float | np.ndarray = 0.1,
    use_genes_column: str | None = "highly_variable",
    transferred_adjusted_basis: str = "X_pca_harmony",
    transferred_primary_basis: str = "X_pca_reference",
    ref_basis_loadings: str = "PCs",
    K: int | None = None,
    reference_primary_basis: str = "X_pca",
) -> None:
    """Symphony query mapping (`symphonypy/tl.py:271-397`) on the GPU.

    Writes `adata_query.obsm[transferred_primary_basis]` (projection onto the reference PCs),
    `obsm[transferred_adjusted_basis]` (after the mixture-of-experts correction) and
    `obsm[transferred_adjusted_basis + "_symphony_R"]` (soft cluster memberships, N x K).

    `sigma` is the soft-assignment temperature and `lamb` the ridge penalty per batch level
    (the reference's docstring swaps the two descriptions, `tl.py:295-298`; the behaviour is kept).
    """
    assert "mean" in adata_ref.var, "Gene expression means are expected to be saved in adata_ref.var"
    assert "std" in adata_ref.var, "Gene expression stds are expected to be saved in adata_ref.var"
    if "harmony" not in adata_ref.uns:
        warnings.warn(
            f"Not found `harmony` object in adata_ref.uns.\n"
            f"Assuming that adata_ref doesn't have any batches, "
            f"and using '{reference_primary_basis}' representation of adata_ref for clustering.\n"
            f"Otherwise, firstly run symphonypy.pp.harmony_integrate "
            f"on adata_ref to account for them."
        )
        _run_soft_kmeans(adata_ref, ref_basis=reference_primary_basis, ref_basis_loadings=ref_basis_loadings,
                         K=K, sigma=sigma)
    if use_genes_column is not None:
        assert use_genes_column in adata_ref.var, \
            f"Column `{use_genes_column}` not found in adata_ref.var. Set `use_genes_column` parameter properly"
    if "log1p" not in adata_query.uns:
        warnings.warn("Gene expressions in adata_query should be log1p-transformed")

    args = (adata_query, adata_ref, key, lamb, sigma, use_genes_column, transferred_adjusted_basis, transferred_primary_basis)
    try:
        _map_embedding_once(*args)
    except _NotCanonical:
        # a CSR query with unsorted or duplicated column indices (rare: scipy and AnnData keep matrices canonical):
        # duplicates must be summed before the scale + clip, as the reference's densification does
        Xc = adata_query.X.tocsr().copy()
        Xc.sum_duplicates()
        _map_embedding_once(*args, X=Xc)


def _map_embedding_once(adata_query, adata_ref, key, lamb, sigma, use_genes_column, transferred_adjusted_basis,
                        transferred_primary_basis, X=None) -> None:
    dev = engine.require_cuda(settings.device)
    harmony_ref = adata_ref.uns["harmony"]
    ref_basis_loadings = harmony_ref["ref_basis_loadings"]  # `tl.py:347` overrides the keyword
    checks: list = []

    with torch.cuda.device(dev):
        # soft-assignment operands first: the assignment of a chunk of cells runs right behind its projection
        Kc = int(harmony_ref["K"])
        Cn = np.asarray(harmony_ref["C"], dtype=np.float64)
        Y = Cn / np.linalg.norm(Cn, ord=2, axis=1, keepdims=True)  # `tl.py:361-362`
        inv_sig = _inv_sigma(sigma, Kc, dev)
        Yd = torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float32)).to(dev)
        N = int(adata_query.X.shape[0])
        R = torch.empty((N, Kc), device=dev, dtype=torch.float32)
        z_host, z_h = _pinned_result((N, Cn.shape[1]))
        r_host, r_h = _pinned_result((N, Kc)) if settings.store_R else (None, None)
        d2h = torch.cuda.Stream(dev)

        def after_chunk(Zd, s, e):
            # 2. soft assignment of rows [s, e) against the reference centroids (per cell: needs nothing global),
            # then Z and R of these rows go back to the host while the next chunk of X is still coming in
            # (PCIe is full duplex: ~0.5 GB of results hidden behind the 8 GB of input at the BASELINE shape)
            engine.assign(Zd[s:e], Yd, inv_sig, out=R[s:e])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            d2h.wait_event(ev)
            with torch.cuda.stream(d2h):
                z_host[s:e].copy_(Zd[s:e], non_blocking=True)
                if r_host is not None:
                    r_host[s:e].copy_(R[s:e], non_blocking=True)

        # 1. project
        Z = _project(adata_query, adata_ref, ref_basis_loadings, use_genes_column, dev, after_chunk, X=X,
                     pending_checks=checks)
        d = Z.shape[1]
        assert Yd.shape == (Kc, d), "harmony C must be [K, n_components]"

        # 3. mixture-of-experts correction
        # (these small synchronous uploads wait for the last chunk's projection, which step 3 needs anyway; issuing
        # them before the query streams in was measured 1 ms slower: they then delay the first chunk instead)
        codes_h, levels = _design.batch_codes(adata_query, key)
        n_levels = [len(lv) for lv in levels]
        lam = _design.ridge_diagonal(lamb, n_levels)
        codes = torch.from_numpy(codes_h).to(dev)
        Nr = torch.from_numpy(np.ascontiguousarray(harmony_ref["Nr"], dtype=np.float64)).to(dev)
        Cd = torch.from_numpy(np.ascontiguousarray(Cn)).to(dev)
        R, Zc, info = pipeline.soft_assign_and_correct(Z, Yd, inv_sig, codes, n_levels, torch.from_numpy(lam).to(dev),
                                                       Nr, Cd, R=R)
        if any(int(f.item()) for f in checks):   # (first sync point of the call)
            raise _NotCanonical()
        pipeline.check_solve(info)

        # 4. results back into the AnnData, as the reference does (`tl.py:394-397`, `_utils.py:306`)
        zc_h = _to_host_async(Zc)
        torch.cuda.current_stream(dev).synchronize()
        d2h.synchronize()

        adata_query.obsm[transferred_primary_basis] = z_h
        adata_query.obsm[transferred_adjusted_basis] = zc_h
        _remember(z_h, Z)
        _remember(zc_h, Zc)
        if r_h is not None:
            adata_query.obsm[f"{transferred_adjusted_basis}_symphony_R"] = r_h
            _remember(r_h, R)

# ----------------------------------------------------------------------------- per-cell confidence
def per_cell_confidence(
    adata_query,
    adata_ref,
    ref_basis_adjusted: str = "X_pca_harmony",
    query_basis_adjusted: str = "X_pca_harmony",
    transferred_primary_basis: str = "X_pca_reference",
    obs: str = "symphony_per_cell_dist",
) -> None:
    """R-weighted Mahalanobis distance of every query cell to the reference clusters
    (`symphonypy/tl.py:33-111`); higher = less confident.  Writes `adata_query.obs[obs]`.

    As in the reference, the inverse cluster covariances and centres of the reference are computed on first
    use (`symphonypy/_utils.py:468-496`) and cached in `adata_ref.uns["harmony"]` under `inv_cluster_covs`
    [K, d, d] and `cluster_centers` [K, d, 1]."""
    assert "harmony" in adata_ref.uns, \
        "Harmony object not found in adata_ref.uns['harmony']. First, run standard symphony query mapping."
    harmony = adata_ref.uns["harmony"]
    dev = engine.require_cuda(settings.device)
    with torch.cuda.device(dev):
        if ("inv_cluster_covs" not in harmony) or ("cluster_centers" not in harmony):
            R_ref = torch.from_numpy(np.ascontiguousarray(harmony["R"], dtype=np.float32)).to(dev)
            X_ref = _device_copy(adata_ref.obsm[ref_basis_adjusted], dev)
            inv_cov, mean = engine.cluster_covariances(X_ref, R_ref)
            harmony["inv_cluster_covs"] = inv_cov.cpu().numpy()
            harmony["cluster_centers"] = mean.cpu().numpy()[..., np.newaxis]
        else:
            logger.info("Using precomputed reference's clusters covariance matrices from the Harmony object")
        inv_cov = torch.from_numpy(np.ascontiguousarray(harmony["inv_cluster_covs"], dtype=np.float32)).to(dev)
        mean = torch.from_numpy(np.ascontiguousarray(np.asarray(harmony["cluster_centers"])[..., 0], dtype=np.float32)).to(dev)
        Xq = _device_copy(adata_query.obsm[transferred_primary_basis], dev)
        Rq = _device_copy(adata_query.obsm[f"{query_basis_adjusted}_symphony_R"], dev)
        out = engine.cell_confidence(Xq, Rq, mean, inv_cov)
        adata_query.obs[obs] = _to_host(out)


# ----------------------------------------------------------------------------- per-cluster confidence
def _cluster_maha_dists(size, mean, scatter, centroids, centroids_norm, u, lamb) -> np.ndarray:
    """`_cluster_maha_dist` (`symphonypy/_utils.py:436-463`) for all clusters at once, from their sizes [G],
    centroids [G, d] and centred scatter matrices sum (x-mean)(x-mean)^T [G, d, d]: Mahalanobis distance to the
    reference centroid with the largest cosine, NaN for clusters with fewer than u*d cells."""
    G, d = mean.shape
    dist = np.full(G, np.nan)
    for g in range(G):
        if size[g] < u * d:
            logger.info("cluster contains too few cells to estimate confidence: ', query_cluster_labels_unique[c])")
            continue
        cov = scatter[g] / (size[g] - 1) + np.diag([lamb] * d)
        closest = centroids[np.argmax(centroids_norm @ mean[g])]
        dif = closest - mean[g]
        dist[g] = float((dif @ np.linalg.inv(cov) @ dif) ** 0.5)
    return dist


def per_cluster_confidence(
    adata_query,
    adata_ref,
    cluster_key,
    u: float = 2,
    lamb: float = 0,
    transferred_primary_basis: str = "X_pca_reference",
    obs: str | None = "symphony_per_cluster_dist",
    uns: str | None = "symphony_per_cluster_dist",
) -> None:
    """Mahalanobis distance from every user-defined query cluster to its nearest reference centroid
    (`symphonypy/tl.py:114-179`, `_utils.py:422-465`); higher = less confident.  Clusters with fewer than
    `u * d` cells get no score (NaN).  Writes `adata_query.uns[uns]` = {key, dist, cluster_labels} and, per cell,
    `adata_query.obs[obs]`.

    The size, centroid and covariance of all clusters come from two passes over the embedding on the GPU
    (`engine.group_moments`); the reference slices the AnnData once per cluster.  The G small d x d systems are
    finished on the host with `np.linalg.inv`, so a singular covariance raises `LinAlgError` as the reference does.

    Under `torch.distributed` the moments are those of THIS rank's shard (no reduction across ranks): call it
    on the whole query.

    Quirk: the reference's per-cell broadcast (`tl.py:179`, `dists[x[0]]`) relies on positional `Series[0]`, which
    raises `KeyError` under pandas >= 3; the intended result (every cell gets its cluster's score) is written."""
    assert "harmony" in adata_ref.uns, \
        "Harmony object not found in adata_ref.uns['harmony']. First, run standard symphony query mapping."
    harmony = adata_ref.uns["harmony"]
    C_ = np.asarray(harmony["C"], dtype=np.float64)
    Nr = np.asarray(harmony["Nr"], dtype=np.float64)
    centroids = C_ / Nr[..., np.newaxis]
    centroids_norm = centroids / np.linalg.norm(centroids, ord=2, axis=1, keepdims=True)

    # observed=True: with categorical keys pandas < 3 would otherwise list the full cartesian product in size().index
    # while ngroup() numbers the observed groups only
    grouped = adata_query.obs.groupby(cluster_key, observed=True, sort=True)
    labels = grouped.size().index
    codes_h = grouped.ngroup().to_numpy()
    G = len(labels)
    assert G > 0 and codes_h.max(initial=-1) < G, "per_cluster_confidence: no clusters found"
    dev = engine.require_cuda(settings.device)
    with torch.cuda.device(dev):
        Xq = _device_copy(adata_query.obsm[transferred_primary_basis], dev)
        d = Xq.shape[1]
        # cells with a missing key (ngroup == -1) go to an extra group that is dropped
        codes = torch.from_numpy(np.where(codes_h < 0, G, codes_h).astype(np.int32)).to(dev)
        size, mean, scatter = engine.group_moments(Xq, codes, G + 1)
        size_h = size[:G].cpu().numpy()
        mean_h = mean[:G].cpu().numpy()
        scatter_h = scatter[:G].cpu().numpy()

    dist = _cluster_maha_dists(size_h, mean_h, scatter_h, centroids, centroids_norm, u, lamb)

    if uns is not None:
        adata_query.uns[uns] = {"key": cluster_key, "dist": dist, "cluster_labels": labels.to_numpy()}
    if obs is not None:
        adata_query.obs[obs] = np.where(codes_h >= 0, dist[np.clip(codes_h, 0, G - 1)], np.nan)


# ----------------------------------------------------------------------------- kNN transfer
_KNN_DEFAULTS = dict(n_neighbors=5, weights="uniform", algorithm="auto", leaf_size=30, p=2, metric="minkowski",
                     metric_params=None, n_jobs=None)
_KNN_ORDER = list(_KNN_DEFAULTS)


def _knn_params(args, kwargs) -> dict:
    """Accepts the `KNeighborsClassifier(*args, **kwargs)` surface of `tl.py:425` and rejects what the
    kernels do not implement instead of silently computing something else."""
    if len(args) > 1:
        # sklearn's estimator takes n_neighbors positionally and everything else by keyword
        raise TypeError(f"KNeighborsClassifier.__init__() takes from 1 to 2 positional arguments but "
                        f"{len(args) + 1} were given")
    p = dict(_KNN_DEFAULTS)
    if args:
        p["n_neighbors"] = args[0]
    for k, v in kwargs.items():
        if k not in p:
            raise TypeError(f"KNeighborsClassifier.__init__() got an unexpected keyword argument '{k}'")
        p[k] = v
    k = p["n_neighbors"]
    if not isinstance(k, numbers.Integral) or isinstance(k, bool) or k < 1:
        raise ValueError(f"The 'n_neighbors' parameter of KNeighborsClassifier must be an int in the range "
                         f"[1, inf) or None. Got {k!r} instead.")
    if p["weights"] not in ("uniform", "distance", None):
        raise NotImplementedError("symphonypy_b200 implements weights='uniform' and 'distance' (no CPU fallback "
                                  "for callables)")
    if p["algorithm"] not in ("auto", "brute", "kd_tree", "ball_tree"):
        raise ValueError(f"The 'algorithm' parameter of KNeighborsClassifier must be a str among "
                         f"{{'auto', 'ball_tree', 'brute', 'kd_tree'}}. Got {p['algorithm']!r} instead.")
    metric, pw = p["metric"], p["p"]
    euclid = metric in ("euclidean", "l2") or (metric == "minkowski" and pw == 2)
    if not euclid or p["metric_params"] is not None:
        raise NotImplementedError("symphonypy_b200 implements the Euclidean metric only (no CPU fallback)")
    return p


def _label_columns(adata_ref, ref_labels):
    y = adata_ref.obs[ref_labels]
    if isinstance(y, pd.DataFrame):
        return [y[c] for c in y.columns], True
    return [y], False


def _encode_labels(col: pd.Series, dev):
    """sklearn's `classes_` (sorted unique labels) and the integer code of every reference cell, on the device.
    Re-done on every call (no cache: an edited label column must never transfer old labels); the dtype-specific
    paths keep that cheap — categorical columns (what AnnData stores) are remapped from their codes in ~3 ms for
    500k cells, Arrow-backed strings factorize in ~8 ms, object columns in ~20 ms."""
    inv = classes = None
    try:
        if isinstance(col.dtype, pd.CategoricalDtype):
            codes = col.cat.codes.to_numpy()
            if (codes < 0).any():
                raise ValueError("missing labels")
            present = np.unique(codes)
            cats = np.asarray(col.cat.categories)[present]
            order = np.argsort(cats, kind="stable")          # sklearn: np.unique(y) order
            lut = np.zeros(len(col.cat.categories), dtype=np.int32)
            lut[present[order]] = np.arange(len(order), dtype=np.int32)
            inv, classes = lut[codes], cats[order]
        else:
            inv, classes = pd.factorize(col, sort=True)
            classes = np.asarray(classes)
            if classes.dtype.kind in "OUT" or isinstance(col.dtype, pd.StringDtype):
                classes = classes.astype(object)
            if (inv < 0).any():
                raise ValueError("missing labels")
    except (TypeError, ValueError):
        classes, inv = np.unique(np.asarray(col), return_inverse=True)
    return classes, torch.from_numpy(np.ascontiguousarray(inv, dtype=np.int32)).to(dev)


_DECODE_CHUNK = 1 << 18
_DECODE_POOL: list = []


def _decode_labels(classes: np.ndarray, lab: np.ndarray, index: pd.Index):
    """`classes[lab]` as the column pandas builds from sklearn's object array of predictions.  When pandas infers
    its Arrow-backed `str` dtype for such an array (pandas >= 3), the same column is produced by an Arrow
    dictionary decode: 15 ms instead of 100 ms for 1M labels (the numpy take + per-element string conversion
    was a third of `transfer_labels_kNN`'s wall time)."""
    if classes.dtype == object:
        dt = pd.Series(classes).dtype
        if isinstance(dt, pd.StringDtype) and getattr(dt, "storage", None) == "pyarrow":
            import pyarrow as pa
            import pyarrow.compute as pc
            names = pa.array(list(classes), type=pa.large_string())
            codes = np.ascontiguousarray(lab, dtype=np.int32)
            if codes.shape[0] < 2 * _DECODE_CHUNK:
                arr = pc.take(names, pa.array(codes))
            else:   # Arrow releases the interpreter lock: the gather of a few hundred thousand strings per thread
                if not _DECODE_POOL:
                    _DECODE_POOL.append(ThreadPoolExecutor(max_workers=4))
                parts = _DECODE_POOL[0].map(lambda o: pc.take(names, pa.array(codes[o:o + _DECODE_CHUNK])),
                                            range(0, codes.shape[0], _DECODE_CHUNK))
                arr = pa.chunked_array(list(parts), type=pa.large_string())
            return pd.Series(pd.array(arr, dtype=dt), index=index)
    return classes[lab]


def transfer_labels_kNN(
    adata_query,
    adata_ref,
    ref_labels: list[str] | str,
    *kNN_args,
    query_labels: list[str] | str | None = None,
    ref_basis: str = "X_pca_harmony",
    query_basis: str = "X_pca_harmony",
    confidence_key: list[str] | str | None = None,
    neighbors_key: str | None = None,
    **kNN_kwargs,
) -> None:
    """kNN label transfer (`symphonypy/tl.py:400-436`) on the GPU.

    `*kNN_args` / `**kNN_kwargs` are the `sklearn.neighbors.KNeighborsClassifier` arguments the
    reference forwards (`tl.py:425`); brute-force Euclidean search with a uniform vote is what every
    accepted combination computes (sklearn picks brute force itself for d > 15).  Ties in the vote go
    to the lowest label in sorted order, distance ties to the lower reference index.

    Extras (not in the reference, off by default): `confidence_key` names `obs` column(s) that receive
    the winning vote fraction (`predict_proba(...).max(1)`, the TODO at `tl.py:432`); `neighbors_key`
    stores neighbour indices / squared distances in `obsm[key + "_indices" / "_sqdist"]`.
    """
    params = _knn_params(kNN_args, kNN_kwargs)
    k = int(params["n_neighbors"])
    dev = engine.require_cuda(settings.device)
    ref_emb = adata_ref.obsm[ref_basis]
    label_cols, two_d = _label_columns(adata_ref, ref_labels)
    M = ref_emb.shape[0]
    if k > M:  # sklearn's message at kneighbors()
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = {M}, "
                         f"n_samples = {adata_query.obsm[query_basis].shape[0]}")
    if query_labels is None:
        query_labels = ref_labels

    with torch.cuda.device(dev):
        # the packed index is found again by the CONTENT of the reference embedding (a replaced or edited obsm array
        # gets a new index)
        # (the digest of the reference embedding runs on a worker thread — a foreign call, no interpreter lock held —
        # while this thread uploads the query, launches the scan and encodes the label columns)
        if not _DIGEST_POOL:
            _DIGEST_POOL.append(ThreadPoolExecutor(max_workers=1))
        digest = _DIGEST_POOL[0].submit(content_digest, ref_emb)
        Q = _device_copy(adata_query.obsm[query_basis], dev)
        encoded: list = []

        def encode_labels():   # host work (pandas factorize: ~15 ms for 500k labels) that overlaps the scan
            if not encoded:
                encoded.extend(_encode_labels(col, dev) for col in label_cols)  # sklearn: classes_ = sorted unique labels

        # The index that was built from this very array object last time is searched SPECULATIVELY while the digest of
        # the array's content is still being computed; the result only stands if the digest then matches the one the
        # index was built from (an in-place edit of the embedding changes it: the search is redone on a fresh index).
        result = None
        spec_key = (id(ref_emb), str(dev)) if isinstance(ref_emb, np.ndarray) else None
        spec = _KNN_SPEC.get(spec_key) if spec_key is not None else None
        if spec is not None and spec[0]() is ref_emb and spec[3] == (ref_emb.shape, ref_emb.dtype.str, ref_emb.ctypes.data):
            result = engine.knn_search_certified(spec[2], Q, k, settings.knn_method, prune=settings.knn_prune,
                                                 after_launch=encode_labels)
            if digest.result() != spec[1]:
                result = None
        if result is None:
            dg = digest.result()
            index = _cached(("knn", str(dev), dg), lambda: engine.knn_fit(_device_copy(ref_emb, dev)))
            if spec_key is not None:
                _KNN_SPEC[spec_key] = (weakref.ref(ref_emb, lambda _r, kk=spec_key: _KNN_SPEC.pop(kk, None)), dg, index,
                                       (ref_emb.shape, ref_emb.dtype.str, ref_emb.ctypes.data))
                while len(_KNN_SPEC) > 4:
                    _KNN_SPEC.popitem(last=False)
            result = engine.knn_search_certified(index, Q, k, settings.knn_method, prune=settings.knn_prune,
                                                 after_launch=encode_labels)
        idx, dist2, n_repaired = result
        encode_labels()
        if n_repaired:
            logger.debug("kNN: %d query rows re-run with the FP32 scan", n_repaired)
        preds, confs = [], []
        for classes, codes in encoded:
            lab, conf = engine.vote(idx, codes, want_conf=confidence_key is not None,
                                    dist2=dist2 if params["weights"] == "distance" else None)
            preds.append((classes, _to_host(lab)))
            if conf is not None:
                confs.append(_to_host(conf))
        if neighbors_key is not None:
            adata_query.obsm[neighbors_key + "_indices"] = _to_host(idx)
            adata_query.obsm[neighbors_key + "_sqdist"] = _to_host(dist2)

    single = query_labels[0] if isinstance(query_labels, list) and len(query_labels) == 1 else query_labels
    if len(preds) == 1 and not isinstance(single, list):
        adata_query.obs[single] = _decode_labels(*preds[0], adata_query.obs.index)
    else:
        cols = [classes[lab] for classes, lab in preds]
        labels_pred = np.stack(cols, axis=1) if two_d and len(cols) > 1 else cols[0]
        if isinstance(query_labels, list) and len(query_labels) == 1:  # `tl.py:434-435`
            labels_pred = labels_pred.reshape(labels_pred.shape[0], 1)
        adata_query.obs[query_labels] = labels_pred
    if confidence_key is not None:
        keys = [confidence_key] if isinstance(confidence_key, str) else list(confidence_key)
        assert len(keys) == len(confs), "confidence_key needs one name per label column"
        for kname, c in zip(keys, confs):
            adata_query.obs[kname] = c
