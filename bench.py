#!/usr/bin/env python
"""Benchmark of the query-mapping hot path (map_embedding + kNN label transfer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--configs LIST]

One "step" = one pass of the whole hot path over one batch of synthetic query cells:
project -> assign -> MoE (stats, all-reduce, solve, apply) -> kNN -> vote.

Headline workload = BASELINE.json configs[1]: 1M query cells onto a 500k-cell reference, 2000 HVGs, 30 PCs,
K=100, k=10 (per GPU; with N GPUs each rank maps its own 1M-cell shard of an N x 1M query against the replicated
reference: weak scaling).  The same JSON line carries, under "configs", short runs of BASELINE configs[2], [3]
and [4] (strong scaling there: the named query is sharded over the ranks), each with its own value / e2e /
roofline / parity, and under "parity" the comparison of the GPU path with the CPU oracle on the very cells the
`cpu_baseline` leg maps.

Output: ONE JSON line (rank 0).  `value` = cells/s with inputs resident in HBM; `e2e` = the same metric through
`symphonypy_b200.tl` with host inputs and host outputs.
"""
from __future__ import annotations

import argparse
import copy
import gc
import json
import logging
import os
import subprocess
import sys
import tempfile
import time
import warnings

# torchrun exports OMP_NUM_THREADS=1 to its workers unless the user set it; rank 0 also times the CPU baseline /
# reference arm, which is to run on all host cores.  OpenBLAS and libgomp size their pools from the environment
# when they are loaded, so the variable has to go before numpy is imported.
if os.environ.get("TORCHELASTIC_RUN_ID") and os.environ.get("OMP_NUM_THREADS") == "1" and \
        os.environ.get("RANK", "0") == "0":
    del os.environ["OMP_NUM_THREADS"]

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N (cells per GPU when scaling is weak, cells of the whole query when strong), M, G, d, K, k, batch levels, cell types
    "config2": dict(N=1_000_000, M=500_000, G=2000, d=30, K=100, k=10, levels=[1], L=20, scaling="weak",
                    desc="BASELINE configs[1]: 1M-cell query onto 500k-cell ref, 2000 HVGs, 30 PCs, K=100, k=10"),
    "config3": dict(N=10_000_000, M=1_000_000, G=2000, d=50, K=200, k=30, levels=[8], L=30, scaling="strong",
                    desc="BASELINE configs[2]: 10M-cell query sharded across the GPUs, 1M-cell ref, 50 PCs, K=200, k=30"),
    "config4": dict(N=2_000_000, M=100_000, G=2000, d=30, K=100, k=10, levels=[100], L=20, scaling="strong",
                    desc="BASELINE configs[3]: MoE-heavy, 2M cells, 100 batch levels, K=100"),
    "config5": dict(N=5_000_000, M=5_000_000, G=2000, d=50, K=100, k=50, levels=[1], L=30, scaling="strong", knn_only=True,
                    desc="BASELINE configs[4]: kNN-heavy, 5M query vs 5M-cell ref embedding, 50 dims, k=50 with "
                         "per-cell confidence (vote fraction)"),
    "config1": dict(N=3000, M=10_000, G=2000, d=20, K=100, k=10, levels=[1], L=12, scaling="weak",
                    desc="BASELINE configs[0]: tutorial shape"),
    "mid": dict(N=200_000, M=500_000, G=2000, d=30, K=100, k=10, levels=[1], L=20, scaling="weak",
                desc="config2 with a 200k-cell query (profiling size)"),
}
METRIC = "query cells mapped/sec (map_embedding + kNN transfer)"
EMB_TOL, LABEL_TOL = 1e-4, 0.999   # north_star: embedding within 1e-4 relative, labels >= 99.9 % identical


def kp_of(d: int, fp16: bool = False) -> int:
    """K extent of the tensor-core operands, padded to 16: BF16x3 = 3 split terms of d values + 3 norm pieces;
    single FP16 = d values + 2 norm pieces."""
    return (d + 2 + 15) // 16 * 16 if fp16 else (3 * d + 3 + 15) // 16 * 16


def knn_mpad(M: int) -> int:
    return (M + 255) // 256 * 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=list(WORKLOADS))
    ap.add_argument("--configs", default="config3,config4,config5",
                    help="other BASELINE configs run briefly after the headline workload ('' or 'none' to skip)")
    ap.add_argument("--config-steps", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and the parity block")
    ap.add_argument("--knn-method", type=int, default=0)
    ap.add_argument("--knn-prune", action="store_true",
                    help="opt-in exact tile pruning of the kNN scan (same results; the headline is the full scan)")
    ap.add_argument("--extras", action="store_true", help="also time the next-row kernels (per_cell_confidence)")
    ap.add_argument("--cpu-n-map", type=int, default=20_000)
    ap.add_argument("--cpu-n-knn", type=int, default=20_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                    power.append(float(p[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples drawing more than half of the observed peak power
        pk = max(power)
        load = [s for s, w in zip(sm, power) if w >= 0.5 * pk] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": pk}


def quiet():
    logging.getLogger("symphonypy").setLevel(logging.ERROR)
    logging.getLogger("symphony_oracle").setLevel(logging.ERROR)


def row_rel_err(a, b) -> float:
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)).max())


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------ CPU oracle sample
def keys_of(w):
    return None if w["levels"] == [1] else [f"batch{v}" for v in range(len(w["levels"]))]


def oracle_sample(w, n_map: int, n_knn: int, seed: int = 0, ref_host=None, want_neighbors: int = 0):
    """The oracle port (numpy + sklearn, the reference's own call sequence) timed on the host cores on a bounded
    sample of the workload: map_embedding on n_map cells and kNN transfer of the first n_knn of them against the
    FULL reference.  Every stage is O(N) at fixed reference size, so cells/s = 1 / (t_map/n_map + t_knn/n_knn).
    Returns (baseline dict, query AnnData as the oracle left it, pristine copy of the query, ref AnnData,
    neighbour indices of the first `want_neighbors` cells or None)."""
    import threadpoolctl
    import torch

    from oracle import symphony_oracle as so
    from symphonypy_b200 import synth

    quiet()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model = synth.make_model(w["G"], w["d"], w["L"], seed, dev)
    Zref, comp, C, Nr = ref_host if ref_host is not None else synth.make_reference(model, w["M"], w["K"], seed)
    n_knn = min(n_knn, n_map)
    if w.get("knn_only"):
        # config 5 names embeddings only: the query embedding comes from the same generator as the reference's
        g = synth._gen(dev, seed + 104)
        qcomp = torch.randint(0, model.L, (n_map,), generator=g, device=dev)
        Zq = model.project(model.expr(qcomp, g))
        X, codes = torch.zeros((n_map, w["G"]), device=dev), torch.zeros((1, n_map), dtype=torch.int32, device=dev)
    else:
        X, codes, _ = synth.make_query(model, n_map, w["levels"], seed + 100)
    query, ref = synth.to_anndata(model, Zref, comp, C, Nr, X, codes, w["K"], seed=seed)
    if w.get("knn_only"):
        query.obsm["X_pca_harmony"] = Zq.double().cpu().numpy()
    pristine = copy.deepcopy(query)
    ncpu = host_threads()
    with warnings.catch_warnings(), threadpoolctl.threadpool_limits(limits=ncpu):
        warnings.simplefilter("ignore")
        t_map = 0.0
        if not w.get("knn_only"):
            t0 = time.perf_counter()
            so.map_embedding(query, ref, key=keys_of(w))
            t_map = time.perf_counter() - t0
        sub = query[np.arange(n_knn), :]  # obsm rows are sliced along
        t0 = time.perf_counter()
        so.transfer_labels_knn(sub, ref, "cell_type", n_neighbors=w["k"])
        t_knn = time.perf_counter() - t0
        info = threadpoolctl.threadpool_info()
        neigh = None
        if want_neighbors:   # not timed: neighbour lists for the index-level parity check
            neigh, _ = so.knn_indices(ref.obsm["X_pca_harmony"], np.asarray(sub.obsm["X_pca_harmony"])[:want_neighbors], w["k"])
    query.obs.loc[query.obs.index[:n_knn], "cell_type"] = np.asarray(sub.obs["cell_type"])
    per_cell = t_map / n_map + t_knn / n_knn
    threads = max([i.get("num_threads", 1) for i in info] + [1])
    import sklearn

    what = (f"kNN transfer of {n_knn} cells against the full {w['M']}-cell reference in {t_knn:.2f}s" if w.get("knn_only") else
            f"map_embedding on {n_map} cells in {t_map:.2f}s + kNN transfer of {n_knn} cells against the full "
            f"{w['M']}-cell reference in {t_knn:.2f}s")
    base = {"value": 1.0 / per_cell, "unit": "cells/s", "cores": int(min(threads, os.cpu_count() or 1)), "kind": "port",
            "sample": f"oracle port (numpy/OpenBLAS + sklearn {sklearn.__version__} KNeighborsClassifier): {what}; "
                      f"extrapolated linearly in N",
            "host_cpus": os.cpu_count(), "t_map_s": t_map, "t_knn_s": t_knn, "n_map": n_map, "n_knn": n_knn,
            "wall_s": t_map + t_knn}
    return base, query, pristine, ref, neigh


def run_reference(args, w):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is pure
    Python and /root/reference does not travel to the GPU box).  One step = map_embedding + kNN transfer of the
    SAME bounded sample of cells (cpu_n_map == cpu_n_knn), so `ms_per_step` is the measured wall time of a step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from symphonypy_b200 import synth

    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model = synth.make_model(w["G"], w["d"], w["L"], 0, dev)
    ref_parts = synth.make_reference(model, w["M"], w["K"], 0)
    n = min(args.cpu_n_map, args.cpu_n_knn)
    walls, base = [], None
    for i in range(args.warmup + args.steps):
        base, *_ = oracle_sample(w, n, n, seed=0, ref_host=ref_parts)
        if i >= args.warmup:
            walls.append(base["wall_s"])
    wall = float(np.mean(walls))
    v = n / wall
    base["value"] = v
    base["sample"] += f" (mean of {args.steps} steps of {n} cells each)"
    # same `config` as our arm prints for this workload (the driver compares them); the bounded sample is described
    # in `sample` / `cpu_baseline`
    n_named = w["N"] if w["scaling"] == "weak" else w["N"] // max(1, args.gpus)
    cfg = dict(workload_config(w, args.workload, n_named, args.gpus, args), numa_node=None)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": base, "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample": {"cells_per_step": n, "note": "bounded CPU sample per step: the same cells go through map_embedding "
                                                "and the kNN transfer; ms_per_step is the measured wall time of a step"},
        "gpu_launches": 0}))


def workload_config(w, name, n_rank, world, args):
    return {"workload": w["desc"], "name": name, "cells_per_gpu": n_rank, "ref_cells": w["M"], "genes": w["G"],
            "pcs": w["d"], "clusters": w["K"], "k": w["k"], "batch_levels": w["levels"],
            "l2": ("inputs larger than L2 (X is %.1f GB per step)" % (n_rank * w["G"] * 4 / 1e9)) if not w.get("knn_only")
                  else "inputs larger than L2 (embeddings + packed reference %.1f GB)" % ((n_rank + 3 * w["M"]) * w["d"] * 4 / 1e9),
            "knn_method": args.knn_method, "knn_prune": bool(args.knn_prune)}


# ------------------------------------------------------------------------------------------ ours
def bind_to_gpu_numa_node(local_rank: int):
    """Multi-rank runs: pin this process (and so its pinned host buffers, which are placed on the allocating
    thread's node) to the NUMA node its GPU hangs off, so that eight ranks streaming 8 GB each do not all pull
    through one socket.  Returns (node, previous affinity) or (None, None); never fails the run."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if not cpus:
            return None, None
        os.sched_setaffinity(0, cpus)
        return node, prev
    except Exception:
        return None, None


class Ctx:
    """Process-wide facts of a bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path)"
        self.dev = torch.device("cuda", self.local)
        torch.cuda.set_device(self.dev)
        self.numa_node, self.full_affinity = bind_to_gpu_numa_node(self.local) if self.world > 1 else (None, None)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            self.peaks = {}

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return float(x)
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def cells_of_rank(w, ctx) -> int:
    """Cells this rank maps: the named N per rank (weak) or its share of the named N (strong)."""
    if w["scaling"] == "weak":
        return w["N"]
    return w["N"] // ctx.world + (1 if ctx.rank < w["N"] % ctx.world else 0)


def make_state(w, ctx):
    """Synthetic data generated on the device (reference replicated: same seed on every rank) and the
    reference-side operands of every stage."""
    import torch

    from symphonypy_b200 import engine, synth, tl

    dev = ctx.dev
    N = cells_of_rank(w, ctx)
    M, G, d, K, levels = w["M"], w["G"], w["d"], w["K"], w["levels"]
    st = {"w": w, "N": N}
    model = synth.make_model(G, d, w["L"], 0, dev)
    Zref, comp, C, Nr = synth.make_reference(model, M, K, 0)
    st.update(model=model, Zref=Zref, comp=comp, C=C, Nr=Nr)
    if w.get("knn_only"):
        g = synth._gen(dev, 2000 + ctx.rank)
        Zq = torch.empty((N, d), device=dev, dtype=torch.float32)
        for s in range(0, N, 32768):
            e = min(N, s + 32768)
            qcomp = torch.randint(0, model.L, (e - s,), generator=g, device=dev)
            Zq[s:e] = model.project(model.expr(qcomp, g))
        st["Zq"] = Zq
    else:
        X, codes, _ = synth.make_query(model, N, levels, 1000 + ctx.rank)
        st.update(X=X, codes=codes)
        st["ld"] = engine.Loadings(mu=model.mean, inv_sd=(1.0 / model.std).contiguous(),
                                   U=torch.nn.functional.pad(model.U, (0, (-d) % 4)).contiguous(), d=d)
        st["Y"] = (C / C.norm(dim=1, keepdim=True)).float().contiguous()
        st["inv_sigma"] = tl._inv_sigma(0.1, K, dev)
        st["lam"] = torch.from_numpy(np.insert(np.ones(sum(levels)), 0, 0.0)).to(dev)
    st["index"] = engine.knn_fit(Zref)
    st["ref_codes"] = comp.int().contiguous()
    torch.cuda.synchronize()
    return st


def run_device(st, ctx, steps: int, warmup: int):
    """K timed steps with inputs resident in HBM.  Returns dict(elapsed_ms, stage_ms, knn_ms, repaired, launches)."""
    import torch

    from symphonypy_b200 import _lib, engine, pipeline

    w, args = st["w"], ctx.args
    stage_ms: dict = {}
    scan_events: list = []
    repaired = [0]
    last = [None]

    def step(record: bool):
        evs = [("start", torch.cuda.Event(enable_timing=True))]
        evs[0][1].record()

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append((name, e))

        if w.get("knn_only"):
            Zc, info = st["Zq"], None
        else:
            Z = engine.project_dense(st["X"], st["ld"])
            mark("project")
            R, Zc, info = pipeline.soft_assign_and_correct(Z, st["Y"], st["inv_sigma"], st["codes"], w["levels"], st["lam"],
                                                           st["Nr"], st["C"], mark)
        lab, conf, idx, dist2, n_rep = pipeline.knn_transfer(st["index"], Zc, w["k"], st["ref_codes"], args.knn_method, mark,
                                                             scan_events if record else None, prune=args.knn_prune)
        if record:
            repaired[0] += n_rep
        last[0] = (lab, conf, info)
        return evs if record else None

    for _ in range(warmup):
        step(False)
    ctx.barrier()
    n0 = _lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    recs = []
    t_start.record()
    for _ in range(steps):
        recs.append(step(True))
    t_end.record()
    ctx.barrier()
    launches = _lib.launch_count() - n0
    elapsed_ms = ctx.max_over_ranks(t_start.elapsed_time(t_end))
    if last[0][2] is not None:
        pipeline.check_solve(last[0][2])
    for evs in recs:
        for (_, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1) / steps
    knn_ms = float(np.mean([a.elapsed_time(b) for a, b in scan_events])) if scan_events else stage_ms.get("knn", float("nan"))
    return {"elapsed_ms": elapsed_ms, "stage_ms": stage_ms, "knn_ms": knn_ms, "repaired": repaired[0] / steps,
            "launches": int(launches)}


def roofline_of(st, ctx, res, name):
    """Roofline of the dominant kernel (the kNN candidate scan) + the memory-bound kernels of the other stages.
    achieved = ALGORITHMIC work per launch / CUDA-event time on the launching stream, measured in this run."""
    from symphonypy_b200 import engine

    w, args, peaks = st["w"], ctx.args, ctx.peaks
    N, M, d, K, G, k = st["N"], w["M"], w["d"], w["K"], w["G"], w["k"]
    V = len(w["levels"])
    knn_ms = res["knn_ms"]
    flops = 2.0 * N * M * d  # algorithmic: one d-term dot product per (query, reference) pair
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = flops / (knn_ms / 1000.0) / 1e12
    # which scan ran as the first tier (engine.knn_search_certified): single-FP16 unless the index fell back
    tier = args.knn_method if args.knn_method else st["index"].first_tier.get(
        k, engine.KNN_TC16 if os.environ.get("SYM_KNN_FP16", "1") != "0" else engine.KNN_TC)
    tier_name = {engine.KNN_FFMA: "knn_scan_ffma_kernel, FP32", engine.KNN_TC: "knn_scan_tc_kernel, BF16x3 tcgen05",
                 engine.KNN_TC16: "knn_scan_tc_kernel, single-FP16 tcgen05 (uncertified rows re-run with BF16x3)"}[tier]
    roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            "traffic": None, "kernel": tier_name + " (+ knn_rerank_kernel)", "ms_per_launch": knn_ms,
            "tensor_flops_issued_per_launch": 2.0 * N * knn_mpad(M) * kp_of(d, tier == engine.KNN_TC16),
            "rows_repaired_per_step": res["repaired"],
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (profiling guide)",
            "algorithmic_flops_per_launch": flops}
    roof["kernel"] = roof["kernel"].replace("project_tc_kernel", "project_tma_kernel")
    if args.knn_prune and st["index"].scan_stats is not None:
        # opt-in tile pruning: the rate is counted on the tiles that were actually scanned
        scanned, full = (int(v) for v in st["index"].scan_stats.cpu())
        roof["pruning"] = {"tiles_scanned": scanned, "tiles_full_scan": full, "scanned_frac": scanned / max(full, 1),
                           "note": "achieved / frac count the flops of the scanned tiles only"}
        roof["achieved"] = ach * scanned / max(full, 1)
        roof["frac"] = roof["achieved"] / peak_tf
        roof["algorithmic_flops_per_launch"] = flops * scanned / max(full, 1)
    try:   # DRAM bytes of one launch: from the committed ncu capture of this workload (not measurable inside the run)
        prof = json.load(open(os.path.join(ROOT, "profiles", "knn_traffic.json"))).get(name, {})
        if prof.get("dram_bytes_per_launch") is not None and ctx.world == 1 and prof.get("cells") in (None, N):
            roof["traffic"] = prof["dram_bytes_per_launch"]
            roof["traffic_source"] = "ncu --set full capture " + str(prof.get("capture", "profiles/"))
    except Exception:
        pass
    if not w.get("knn_only"):
        hbm = peaks.get("hbm_gbs", 6650.0)
        sm = res["stage_ms"]
        other = []
        # FP32 FFMA peak (nominal: 148 SMs x 128 lanes x 2 flop x the SM clock; no measured figure in MEASURED_PEAKS.json):
        # the three K x d products per cell sit at the ridge of the two rooflines (11.5 flop/B at d=30, K=100), so the
        # FP32 fraction is reported next to the HBM one
        fp32_tf = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        for stage, kern, bytes_cell, flops_cell in (
                ("project", "project_tma_kernel (tcgen05 FP16x3, X tiles by TMA)", 4.0 * G + 4.0 * d, None),
                ("assign", "assign_kernel", 4.0 * d + 4.0 * K, 2.0 * K * d),
                ("moe_stats", "moe_stats_kernel (+ counting sort)", 4.0 * d + 4.0 * K + 4.0 * V, 2.0 * K * (d + 1)),
                ("moe_solve_apply", "moe_solve_kernel + moe_apply_kernel", 8.0 * d + 4.0 * K + 4.0 * V, 2.0 * K * d)):
            if stage in sm and sm[stage] > 0:
                gbs = N * bytes_cell / (sm[stage] / 1000.0) / 1e9
                row = {"stage": stage, "kernel": kern, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                       "frac": gbs / hbm, "ms": sm[stage], "algorithmic_bytes_per_cell": bytes_cell}
                if flops_cell:
                    tf = N * flops_cell / (sm[stage] / 1000.0) / 1e12
                    row["fp32"] = {"achieved_tflops": tf, "peak_tflops_nominal": fp32_tf, "frac": tf / fp32_tf,
                                   "algorithmic_flops_per_cell": flops_cell}
                other.append(row)
        roof["other_kernels"] = other
    return roof


def host_query(st, ctx, n_cells: int, fmt: str):
    """The first n_cells of this rank's query as the AnnData a user would hold: X in host memory as a pinned dense
    array, a pageable dense array or a scipy CSR matrix; batch keys as categorical obs columns."""
    import pandas as pd
    import torch

    from symphonypy_b200 import synth
    from symphonypy_b200.anndata_lite import AnnDataLite

    w = st["w"]
    G, levels = w["G"], w["levels"]
    X = st["X"][:n_cells]
    if fmt in ("csr", "csr_pinned"):
        import scipy.sparse as sp

        nz = X != 0
        indptr = torch.zeros(n_cells + 1, dtype=torch.int64, device=ctx.dev)
        indptr[1:] = nz.sum(1).cumsum(0)
        cols = nz.nonzero()[:, 1].int()
        vals = X[nz]
        if fmt == "csr_pinned":   # the matrix's own arrays live in page-locked memory (scipy keeps them as they are)
            dh = torch.empty(vals.numel(), dtype=torch.float32, pin_memory=True)
            ih = torch.empty(cols.numel(), dtype=torch.int32, pin_memory=True)
            dh.copy_(vals)
            ih.copy_(cols)
            Xh = sp.csr_matrix((dh.numpy(), ih.numpy(), indptr.cpu().numpy()), shape=(n_cells, G))
            assert torch.from_numpy(Xh.data).is_pinned() and torch.from_numpy(Xh.indices).is_pinned()
        else:
            Xh = sp.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), indptr.cpu().numpy()), shape=(n_cells, G))
        del vals
        h2d = int(Xh.nnz) * 8 + (n_cells + 1) * 8
        del nz, cols
    else:
        t = torch.empty((n_cells, G), dtype=torch.float32, pin_memory=(fmt == "pinned"))
        t.copy_(X)
        Xh = t.numpy()
        h2d = n_cells * G * 4
    obs = pd.DataFrame(index=pd.RangeIndex(n_cells).astype(str))
    if levels != [1]:
        ch = st["codes"][:, :n_cells].cpu().numpy()
        for v, lv in enumerate(levels):
            obs[f"batch{v}"] = pd.Categorical.from_codes(ch[v], synth.level_names(v, lv))
    q = AnnDataLite(Xh, obs=obs, var=pd.DataFrame(index=pd.Index([f"g{i}" for i in range(G)])), uns={"log1p": {}})
    return q, h2d + n_cells * 4 * (len(levels) if levels != [1] else 0)


def ref_anndata(st):
    from symphonypy_b200 import synth

    w = st["w"]
    if w.get("knn_only"):
        import torch
        X8, c8 = torch.zeros((8, w["G"]), device=st["Zref"].device), torch.zeros((1, 8), dtype=torch.int32, device=st["Zref"].device)
    else:
        X8, c8 = st["X"][:8], st["codes"][:, :8]
    _, ref = synth.to_anndata(st["model"], st["Zref"], st["comp"], st["C"], st["Nr"], X8, c8, w["K"], seed=0)
    return ref


def time_api(ctx, fn, warm: int, steps: int, before=None):
    import torch

    times = []
    for i in range(warm + steps):
        if before is not None:
            before()
        ctx.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return ctx.max_over_ranks(float(np.mean(times)))


def run_e2e(st, ctx, steps: int, variants=("dense_pinned", "dense_pageable", "csr_pinned", "csr_host", "cold"), n_cells=None):
    """End to end through the public API: host inputs -> host outputs, at this rank's cells (or the first n_cells).
    dense_pinned: X is a page-locked float32 array (the headline);  dense_pageable: an ordinary numpy array;
    csr_pinned: a scipy CSR matrix whose arrays are page-locked (the headline);  csr_host: the same in pageable memory
    (what AnnData.X usually is);  cold: csr_host right after tl.clear_caches(),
    i.e. including the reference-side work the reference itself redoes on every call (knn.fit, label encoding)."""
    from symphonypy_b200 import synth, tl

    w = st["w"]
    N = st["N"] if n_cells is None else min(n_cells, st["N"])
    d, K, k = w["d"], w["K"], w["k"]
    ref = ref_anndata(st)
    key = keys_of(w)
    names = set(synth.type_names(w["L"]))
    out = {}

    if w.get("knn_only"):
        import pandas as pd
        from symphonypy_b200.anndata_lite import AnnDataLite

        q = AnnDataLite(np.zeros((N, 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(N).astype(str)),
                        obsm={"X_pca_harmony": st["Zq"][:N].cpu().numpy()})

        def call():
            tl.transfer_labels_kNN(q, ref, "cell_type", n_neighbors=k, confidence_key="cell_type_confidence")

        for name, warm, nst, before in (("dense_pageable", 1, steps, None), ("cold", 0, 1, tl.clear_caches)):
            if name not in variants:
                continue
            dt = time_api(ctx, call, warm, nst, before)
            out[name] = {"value": ctx.world * N / dt, "unit": "cells/s", "ms_per_step": dt * 1000.0,
                         "h2d_bytes_per_step": N * d * 4, "d2h_bytes_per_step": N * 8, "steps": nst, "warmup": warm}
        assert set(np.unique(np.asarray(q.obs["cell_type"]))) <= names
        out["api"] = "symphonypy_b200.tl.transfer_labels_kNN(confidence_key=...), numpy float32 embeddings in, obs columns out"
        out["cells_per_gpu"] = N
        head = out.get("dense_pageable") or next(iter(out.values()))
        return dict(head, **{"variants": out})

    d2h = N * d * 4 * 2 + N * K * 4 + N * 4
    for name, fmt, warm, nst, before in (("dense_pinned", "pinned", 3, steps, None),
                                         ("dense_pageable", "pageable", 1, max(1, steps - 1), None),
                                         ("csr_pinned", "csr_pinned", 3, steps, None),
                                         ("csr_host", "csr", 3, steps, None),
                                         ("cold", "csr", 0, 1, tl.clear_caches)):
        if name not in variants:
            continue
        q, h2d = host_query(st, ctx, N, fmt)

        def call():
            tl.map_embedding(q, ref, key=key)
            tl.transfer_labels_kNN(q, ref, "cell_type", n_neighbors=k)

        dt = time_api(ctx, call, warm, nst, before)
        assert set(np.unique(np.asarray(q.obs["cell_type"]))) <= names
        out[name] = {"value": ctx.world * N / dt, "unit": "cells/s", "ms_per_step": dt * 1000.0, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "steps": nst, "warmup": warm}
        del q
        gc.collect()
    # headline: X as a scipy CSR float32 matrix — what AnnData.X is in practice ("CSR f32 typical") — whose
    # arrays lie in page-locked host memory (the bench contract's "inputs in pinned host memory", as dense_pinned is for a
    # dense array); csr_host is the same matrix in ordinary pageable memory (staged by the host's threads), the dense
    # variants are what the same call costs when X is a dense array
    first = next((n for n in ("csr_pinned", "csr_host", "dense_pinned") if n in out), next(iter(out)))
    head = dict(out[first])
    head["input"] = first
    head["api"] = "symphonypy_b200.tl.map_embedding + tl.transfer_labels_kNN, numpy / scipy X in host memory in, numpy obsm/obs out"
    head["cells_per_gpu"] = N
    head["variants"] = out
    return head


def run_parity(w, ctx, st, n_map: int, n_knn: int, n_neigh: int = 2000):
    """The GPU path against the CPU oracle on the SAME cells: the oracle maps a seeded sample of the workload
    (timed: this is the cpu_baseline leg) and the product API maps the identical AnnData — sharded over the ranks
    when there are several, so the all-reduced MoE statistics are part of what is checked.  Fails the run when the
    embedding error exceeds 1e-4 or labels fall below 99.9 %."""
    import torch
    import torch.distributed as dist

    from symphonypy_b200 import tl

    world, rank = ctx.world, ctx.rank
    n_map -= n_map % world
    n_knn = min(n_knn, n_map)
    base = oq = ref = neigh = None
    ref_host = (st["Zref"], st["comp"], st["C"], st["Nr"])
    if rank == 0:
        if ctx.full_affinity is not None:   # the CPU baseline runs on all host cores, not on this GPU's NUMA node only
            os.sched_setaffinity(0, ctx.full_affinity)
        base, oq, pristine, ref, neigh = oracle_sample(w, n_map, n_knn, ref_host=ref_host, want_neighbors=min(n_neigh, n_knn))
    else:   # same seeded sample and reference objects, no oracle run
        from symphonypy_b200 import synth
        model = st["model"]
        if w.get("knn_only"):
            g = synth._gen(ctx.dev, 104)
            qcomp = torch.randint(0, model.L, (n_map,), generator=g, device=ctx.dev)
            Zq = model.project(model.expr(qcomp, g))
            X, codes = torch.zeros((n_map, w["G"]), device=ctx.dev), torch.zeros((1, n_map), dtype=torch.int32, device=ctx.dev)
        else:
            X, codes, _ = synth.make_query(model, n_map, w["levels"], 100)
        pristine, ref = synth.to_anndata(model, *ref_host, X, codes, w["K"], seed=0)
        if w.get("knn_only"):
            pristine.obsm["X_pca_harmony"] = Zq.double().cpu().numpy()
    ctx.barrier()
    lo, hi = rank * n_map // world, (rank + 1) * n_map // world
    shard = pristine[np.arange(lo, hi), :]
    tl.clear_caches()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if not w.get("knn_only"):
            tl.map_embedding(shard, ref, key=keys_of(w))
        tl.transfer_labels_kNN(shard, ref, "cell_type", n_neighbors=w["k"], neighbors_key="nn")
    d, K = w["d"], w["K"]

    def gathered(a, cols, dtype=torch.float32):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(ctx.dev, dtype).reshape(hi - lo, cols)
        if world == 1:
            return t.cpu().numpy()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return torch.cat(parts).cpu().numpy()

    lab_codes = {n: i for i, n in enumerate(sorted(set(np.asarray(ref.obs["cell_type"], dtype=str))))}
    got_lab = gathered(np.array([lab_codes[x] for x in np.asarray(shard.obs["cell_type"], dtype=str)]), 1, torch.int32)[:, 0]
    got_nn = gathered(shard.obsm["nn_indices"], w["k"], torch.int32)
    got = {}
    if not w.get("knn_only"):
        got["proj"] = gathered(shard.obsm["X_pca_reference"], d)
        got["corr"] = gathered(shard.obsm["X_pca_harmony"], d)
        got["R"] = gathered(shard.obsm["X_pca_harmony_symphony_R"], K)
    if rank != 0:
        return None, None
    par = {"cells": n_map, "knn_cells": n_knn, "sharded_over_ranks": world,
           "against": "oracle port on the same seeded cells (the cpu_baseline sample)"}
    if not w.get("knn_only"):
        par["proj_rel"] = row_rel_err(got["proj"], oq.obsm["X_pca_reference"])
        par["corr_rel"] = row_rel_err(got["corr"], oq.obsm["X_pca_harmony"])
        par["R_abs"] = float(np.abs(got["R"] - oq.obsm["X_pca_harmony_symphony_R"]).max())
    want_lab = np.array([lab_codes[x] for x in np.asarray(oq.obs["cell_type"], dtype=str)[:n_knn]])
    # with a sharded query the kNN runs on the GPU's own corrected embedding (1e-6 from the oracle's): labels on all kNN cells
    par["label_identity"] = float((got_lab[:n_knn] == want_lab).mean())
    if neigh is not None:
        m = neigh.shape[0]
        same_set = [set(a) == set(b) for a, b in zip(got_nn[:m].tolist(), neigh.tolist())]
        par["neighbour_set_identity"] = float(np.mean(same_set))
        par["neighbour_cells"] = m
    par["tolerances"] = {"embedding_rel": EMB_TOL, "label_identity": LABEL_TOL}
    ok = par["label_identity"] >= LABEL_TOL and par.get("proj_rel", 0.0) < EMB_TOL and par.get("corr_rel", 0.0) < EMB_TOL
    par["ok"] = bool(ok)
    return base, par


def free_state(st):
    import torch

    st.clear()
    gc.collect()
    torch.cuda.empty_cache()


def run_workload(name, ctx, steps, warmup, e2e_steps, e2e_cells=None, e2e_variants=None, cpu_n_map=None, cpu_n_knn=None):
    """One workload end to end: device-resident value, roofline, e2e, cpu baseline + parity.  Returns the dict of
    the JSON line's workload-specific keys."""
    from symphonypy_b200 import tl

    args = ctx.args
    w = WORKLOADS[name]
    st = make_state(w, ctx)
    res = run_device(st, ctx, steps, warmup)
    N = st["N"]
    total_cells = ctx.world * N if w["scaling"] == "weak" else w["N"]
    out = {"value": total_cells * steps / (res["elapsed_ms"] / 1000.0), "unit": "cells/s", "steps": steps, "warmup": warmup,
           "ms_per_step": res["elapsed_ms"] / steps, "scaling": w["scaling"],
           "config": dict(workload_config(w, name, N, ctx.world, args), numa_node=ctx.numa_node),
           "stage_ms": res["stage_ms"], "roofline": roofline_of(st, ctx, res, name), "gpu_launches": res["launches"]}
    if not args.no_e2e:
        kw = {} if e2e_variants is None else {"variants": e2e_variants}
        out["e2e"] = run_e2e(st, ctx, e2e_steps, n_cells=e2e_cells, **kw)
        if e2e_cells is not None and e2e_cells < N:
            out["e2e"]["note"] = f"measured on the first {min(e2e_cells, N)} cells of each rank's shard"
            out["e2e"]["value"] = out["e2e"]["value"]   # cells/s of the measured slice, whole job
    tl.clear_caches()
    if not args.no_cpu:
        base, par = run_parity(w, ctx, st, cpu_n_map or args.cpu_n_map, cpu_n_knn or args.cpu_n_knn)
        out["cpu_baseline"], out["parity"] = base, par
    else:
        out["cpu_baseline"], out["parity"] = None, None
    tl.clear_caches()
    return out, st


def main():
    args = parse()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist

    from symphonypy_b200 import _lib, engine, tl

    ctx = Ctx(args)
    tl.settings.device = ctx.dev
    tl.settings.knn_method = args.knn_method
    tl.settings.knn_prune = bool(args.knn_prune)
    quiet()
    _lib.load()
    rank = ctx.rank

    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    head, st = run_workload(args.workload, ctx, args.steps, args.warmup, args.e2e_steps)
    clocks = sampler.stop() if rank == 0 else None

    extras = None
    if args.extras and rank == 0 and not w.get("knn_only"):
        from symphonypy_b200 import synth

        # next row: R-weighted Mahalanobis confidence, device resident
        M, K, d, N = w["M"], w["K"], w["d"], st["N"]
        Zr2, _, _, _, R_ref = synth.make_reference(st["model"], M, K, 0, keep_R=True)
        t0e, t1e, t2e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        Zq = engine.project_dense(st["X"], st["ld"])
        Rq = engine.assign(Zq, st["Y"], st["inv_sigma"])
        Rk = R_ref.T.float().contiguous()
        torch.cuda.synchronize()
        t0e.record()
        inv_cov, mean = engine.cluster_covariances(Zr2, Rk)
        t1e.record()
        for _ in range(3):
            engine.cell_confidence(Zq, Rq, mean, inv_cov)
        t2e.record()
        torch.cuda.synchronize()
        extras = {"per_cell_confidence": {"reference_covariances_ms": t0e.elapsed_time(t1e),
                                          "query_ms": t1e.elapsed_time(t2e) / 3.0, "cells": N, "ref_cells": M,
                                          "algorithmic_flops_query": 2.0 * N * K * (d * d + d),
                                          "tflops_query": 2.0 * N * K * (d * d + d) / (t1e.elapsed_time(t2e) / 3.0) / 1e9}}
        del Zr2, R_ref, Rk, Zq, Rq
    free_state(st)

    # ---- the other BASELINE configs, briefly (strong scaling: the named query is sharded over the ranks)
    others = {}
    names = [c for c in args.configs.split(",") if c and c != "none" and c != args.workload]
    for name in names:
        wc = WORKLOADS[name]
        try:
            # e2e on at most 1.25M cells per rank (config 3's 8-GPU shard): a 10M-cell dense X would need 80 GB of
            # page-locked host memory per rank
            o, st_c = run_workload(name, ctx, args.config_steps, 1, 2, e2e_cells=1_250_000,
                                   e2e_variants=("dense_pinned", "csr_pinned", "csr_host") if not wc.get("knn_only") else ("dense_pageable",),
                                   cpu_n_map=10_000 if not wc.get("knn_only") else 2_000,
                                   cpu_n_knn=5_000 if not wc.get("knn_only") else 2_000)
            free_state(st_c)
            others[name] = o
        except Exception as exc:   # a failing side config must not take the headline number with it
            others[name] = {"error": f"{type(exc).__name__}: {exc}"[:500]}
            gc.collect()
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": "cells/s", "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": head["config"],
            "stage_ms": head["stage_ms"], "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"],
            "parity": head["parity"], "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"], "clocks": clocks,
        }
        if others:
            line["configs"] = others
        if extras is not None:
            line["extras"] = extras
        print(json.dumps(line))
    failed = []
    if rank == 0:
        for nm, blk in [(args.workload, head)] + list(others.items()):
            p = blk.get("parity") if isinstance(blk, dict) else None
            if p is not None and not p["ok"]:
                failed.append((nm, p))
    if ctx.world > 1:
        dist.destroy_process_group()
    if failed:
        print("PARITY FAILED: " + json.dumps(failed), file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
