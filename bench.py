#!/usr/bin/env python
"""Benchmark of the query-mapping hot path (map_embedding + kNN label transfer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the whole hot path over one batch of synthetic query cells:
project -> assign -> MoE (stats, all-reduce, solve, apply) -> kNN -> vote.
Default workload is BASELINE.json configs[1]: 1M query cells onto a 500k-cell reference,
2000 HVGs, 30 PCs, K=100, k=10 (per GPU; with N GPUs each rank maps its own 1M-cell shard of an
N x 1M query against the replicated reference: weak scaling).

Output: ONE JSON line (rank 0).  `value` = cells/s with inputs resident in HBM; `e2e` = the
same metric through `symphonypy_b200.tl` with host (pinned) inputs and host outputs.
"""
from __future__ import annotations

import argparse
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
    # name: N per GPU, M, G, d, K, k, batch levels, cell types
    "config2": dict(N=1_000_000, M=500_000, G=2000, d=30, K=100, k=10, levels=[1], L=20,
                    desc="BASELINE configs[1]: 1M-cell query onto 500k-cell ref, 2000 HVGs, 30 PCs, K=100, k=10"),
    "config3": dict(N=1_250_000, M=1_000_000, G=2000, d=50, K=200, k=30, levels=[8], L=30,
                    desc="BASELINE configs[2] per-GPU shard: 1.25M of 10M cells, 1M-cell ref, 50 PCs, K=200, k=30"),
    "config4": dict(N=2_000_000, M=100_000, G=2000, d=30, K=100, k=10, levels=[100], L=20,
                    desc="BASELINE configs[3]: MoE-heavy, 2M cells, 100 batch levels, K=100"),
    "config1": dict(N=3000, M=10_000, G=2000, d=20, K=100, k=10, levels=[1], L=12,
                    desc="BASELINE configs[0]: tutorial shape"),
    "mid": dict(N=200_000, M=500_000, G=2000, d=30, K=100, k=10, levels=[1], L=20,
                desc="config2 with a 200k-cell query (profiling size)"),
}
METRIC = "query cells mapped/sec (map_embedding + kNN transfer)"


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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--knn-method", type=int, default=0)
    ap.add_argument("--extras", action="store_true", help="also time the next-row kernels (per_cell_confidence)")
    ap.add_argument("--cpu-n-map", type=int, default=20_000)
    ap.add_argument("--cpu-n-knn", type=int, default=5_000)
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


# ------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(w, n_map: int, n_knn: int, seed: int = 0, ref_host=None):
    """The oracle port (numpy + sklearn, the reference's own call sequence) timed on the host cores on a
    bounded sample: map_embedding on n_map cells, kNN transfer on n_knn cells against the FULL reference;
    every stage is O(N) at fixed reference size, so cells/s = 1 / (t_map/n_map + t_knn/n_knn)."""
    import threadpoolctl
    import torch

    from oracle import symphony_oracle as so
    from symphonypy_b200 import synth

    logging.getLogger("symphony_oracle").setLevel(logging.ERROR)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model = synth.make_model(w["G"], w["d"], w["L"], seed, dev)
    if ref_host is None:
        Zref, comp, C, Nr = synth.make_reference(model, w["M"], w["K"], seed)
    else:
        Zref, comp, C, Nr = ref_host
    X, codes, _ = synth.make_query(model, n_map, w["levels"], seed + 100)
    query, ref = synth.to_anndata(model, Zref, comp, C, Nr, X, codes, w["K"], seed=seed)
    key = None if w["levels"] == [1] else [f"batch{v}" for v in range(len(w["levels"]))]
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # otherwise pin OpenBLAS and sklearn's OpenMP loops to one core
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with warnings.catch_warnings(), threadpoolctl.threadpool_limits(limits=ncpu):
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        so.map_embedding(query, ref, key=key)
        t_map = time.perf_counter() - t0
        sub = query[np.arange(min(n_knn, n_map)), :]  # obsm rows are sliced along
        t0 = time.perf_counter()
        so.transfer_labels_knn(sub, ref, "cell_type", n_neighbors=w["k"])
        t_knn = time.perf_counter() - t0
        info = threadpoolctl.threadpool_info()
    n_knn = min(n_knn, n_map)
    per_cell = t_map / n_map + t_knn / n_knn
    threads = max([i.get("num_threads", 1) for i in info] + [1])
    return {"value": 1.0 / per_cell, "unit": "cells/s", "cores": int(min(threads, os.cpu_count() or 1)), "kind": "port",
            "sample": f"oracle port (numpy/OpenBLAS + sklearn {__import__('sklearn').__version__} KNeighborsClassifier): "
                      f"map_embedding on {n_map} cells in {t_map:.2f}s + kNN transfer of {n_knn} cells against the "
                      f"full {w['M']}-cell reference in {t_knn:.2f}s; extrapolated linearly in N",
            "host_cpus": os.cpu_count(), "t_map_s": t_map, "t_knn_s": t_knn}


def run_reference(args, w):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is pure
    Python and /root/reference does not travel to the GPU box) on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from symphonypy_b200 import synth

    dev = "cuda" if torch.cuda.is_available() else "cpu"
    model = synth.make_model(w["G"], w["d"], w["L"], 0, dev)
    ref_parts = synth.make_reference(model, w["M"], w["K"], 0)
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(w, args.cpu_n_map, args.cpu_n_knn, seed=0, ref_host=ref_parts)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    n = args.cpu_n_map
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * n / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "cells_per_step": n, "note": "bounded CPU sample per step"},
        "cpu_baseline": base, "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


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


def main():
    args = parse()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist

    from symphonypy_b200 import _lib, engine, pipeline, synth, tl
    from symphonypy_b200.anndata_lite import AnnDataLite

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa_node, full_affinity = bind_to_gpu_numa_node(local) if world > 1 else (None, None)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tl.settings.device = dev
    tl.settings.knn_method = args.knn_method
    logging.getLogger("symphonypy").setLevel(logging.ERROR)
    _lib.load()

    # ---- synthetic data, generated on the device (reference replicated: same seed on every rank)
    N, M, G, d, K, k, levels = w["N"], w["M"], w["G"], w["d"], w["K"], w["k"], w["levels"]
    model = synth.make_model(G, d, w["L"], 0, dev)
    Zref, comp, C, Nr = synth.make_reference(model, M, K, 0)
    X, codes, _ = synth.make_query(model, N, levels, 1000 + rank)
    ld = engine.Loadings(mu=model.mean, inv_sd=(1.0 / model.std).contiguous(),
                         U=torch.nn.functional.pad(model.U, (0, (-d) % 4)).contiguous(), d=d)
    Y = (C / C.norm(dim=1, keepdim=True)).float().contiguous()
    inv_sigma = tl._inv_sigma(0.1, K, dev)
    lam = torch.from_numpy(np.insert(np.ones(sum(levels)), 0, 0.0)).to(dev)
    index = engine.knn_fit(Zref)
    ref_codes = comp.int().contiguous()
    torch.cuda.synchronize()

    stage_ms: dict = {}
    scan_events: list = []
    repaired = [0]

    def step(record: bool):
        evs = [("start", torch.cuda.Event(enable_timing=True))]
        evs[0][1].record()

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append((name, e))

        Z = engine.project_dense(X, ld)
        mark("project")
        R, Zc, info = pipeline.soft_assign_and_correct(Z, Y, inv_sigma, codes, levels, lam, Nr, C, mark)
        lab, conf, idx, dist2, n_rep = pipeline.knn_transfer(index, Zc, k, ref_codes, args.knn_method, mark,
                                                             scan_events if record else None)
        if record:
            repaired[0] += n_rep
        return evs if record else None, (lab, info, n_rep)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    recs = []
    t_start.record()
    for _ in range(args.steps):
        evs, out = step(True)
        recs.append(evs)
    t_end.record()
    barrier()
    launches = _lib.launch_count() - n0
    elapsed_ms = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if rank == 0 else None
    pipeline.check_solve(out[1])
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    for evs in recs:
        for (n0_, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1) / args.steps
    value = world * N * args.steps / (elapsed_ms / 1000.0)

    # ---- roofline of the dominant kernel (the kNN candidate scan; rerank included in the 'knn' stage)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # dominant kernel: the kNN candidate scan (sym_knn = scan + float64 re-rank launches, CUDA events on the stream)
    knn_ms = float(np.mean([a.elapsed_time(b) for a, b in scan_events])) if scan_events else stage_ms.get("knn", float("nan"))
    flops = 2.0 * N * M * d  # algorithmic: one d-term dot product per (query, reference) pair
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = flops / (knn_ms / 1000.0) / 1e12
    # which scan ran as the first tier (engine.knn_search_certified): single-FP16 unless the index fell back
    tier = args.knn_method if args.knn_method else index.first_tier.get(
        k, engine.KNN_TC16 if os.environ.get("SYM_KNN_FP16", "1") != "0" else engine.KNN_TC)
    tier_name = {engine.KNN_FFMA: "knn_scan_ffma_kernel, FP32", engine.KNN_TC: "knn_scan_tc_kernel, BF16x3 tcgen05",
                 engine.KNN_TC16: "knn_scan_tc_kernel, single-FP16 tcgen05 (uncertified rows re-run with BF16x3)"}[tier]
    roofline = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": None, "kernel": tier_name + " (+ knn_rerank_kernel)", "ms_per_launch": knn_ms,
                "tensor_flops_issued_per_launch": 2.0 * N * knn_mpad(M) * kp_of(d, tier == engine.KNN_TC16),
                "rows_repaired_per_step": repaired[0] / args.steps,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (profiling guide)",
                "fp32_ffma_peak_tflops": 74.4, "frac_of_ffma_peak": ach / 74.4,
                "algorithmic_flops_per_launch": flops}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "knn_traffic.json")))
        roofline["traffic"] = prof.get(args.workload, {}).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- end to end through the public API: host (pinned) inputs -> host outputs
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((N, G), dtype=torch.float32).pin_memory()
        Xh.copy_(X)
        import pandas as pd

        obs = pd.DataFrame(index=pd.RangeIndex(N).astype(str))
        key = None
        if levels != [1]:
            key = []
            ch = codes.cpu().numpy()
            for v, lv in enumerate(levels):
                obs[f"batch{v}"] = pd.Categorical.from_codes(ch[v], synth.level_names(v, lv))
                key.append(f"batch{v}")
        q = AnnDataLite(Xh.numpy(), obs=obs, var=pd.DataFrame(index=pd.Index([f"g{i}" for i in range(G)])),
                        uns={"log1p": {}})
        query_, ref = synth.to_anndata(model, Zref, comp, C, Nr, X[:8], codes[:, :8], K, seed=0)
        del query_
        names = synth.type_names(w["L"])
        h2d = N * G * 4 + N * 4 * len(levels)
        d2h = N * d * 4 * 2 + N * K * 4 + N * 4
        times = []
        E2E_WARMUP = 3  # builds the reference-side caches and both generations of pinned result buffers
        for i in range(E2E_WARMUP + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            tl.map_embedding(q, ref, key=key)
            tl.transfer_labels_kNN(q, ref, "cell_type", n_neighbors=k)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= E2E_WARMUP:
                times.append(dt)
        dt = float(np.mean(times))
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert set(np.unique(np.asarray(q.obs["cell_type"]))) <= set(names)
        e2e = {"value": world * N / dt, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1000.0, "steps": args.e2e_steps, "warmup": E2E_WARMUP,
               "api": "symphonypy_b200.tl.map_embedding + tl.transfer_labels_kNN, pinned numpy X in, numpy obsm/obs out"}
        del Xh, q

    extras = None
    if args.extras and rank == 0:
        # next row: R-weighted Mahalanobis confidence, device resident
        Zr2, _, _, _, R_ref = synth.make_reference(model, M, K, 0, keep_R=True)
        t0e, t1e, t2e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        Zq = engine.project_dense(X, ld)
        Rq = engine.assign(Zq, Y, inv_sigma)
        Rk = R_ref.T.float().contiguous()
        torch.cuda.synchronize()
        t0e.record()
        inv_cov, mean = engine.cluster_covariances(Zr2, Rk)
        t1e.record()
        for _ in range(3):
            conf = engine.cell_confidence(Zq, Rq, mean, inv_cov)
        t2e.record()
        torch.cuda.synchronize()
        extras = {"per_cell_confidence": {"reference_covariances_ms": t0e.elapsed_time(t1e),
                                          "query_ms": t1e.elapsed_time(t2e) / 3.0, "cells": N, "ref_cells": M,
                                          "algorithmic_flops_query": 2.0 * N * K * (d * d + d),
                                          "tflops_query": 2.0 * N * K * (d * d + d) / (t1e.elapsed_time(t2e) / 3.0) / 1e9}}
        del Zr2, R_ref, Rk
        # the same query as a scipy CSR matrix in host memory (the usual AnnData.X container): map_embedding end to end
        import pandas as pd
        import scipy.sparse as sp
        nz = X != 0
        indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        indptr[1:] = nz.sum(1).cumsum(0)
        cols = nz.nonzero()[:, 1].int()
        csr = sp.csr_matrix((X[nz].cpu().numpy(), cols.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
        del nz, cols
        qs = AnnDataLite(csr, obs=pd.DataFrame(index=pd.RangeIndex(N).astype(str)),
                         var=pd.DataFrame(index=pd.Index([f"g{i}" for i in range(G)])), uns={"log1p": {}})
        _, ref_s = synth.to_anndata(model, Zref, comp, C, Nr, X[:8], codes[:, :8], K, seed=0)
        ts = []
        for i in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tl.map_embedding(qs, ref_s, key=None)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        extras["map_embedding_csr_host"] = {"ms": 1000.0 * float(np.mean(ts[2:])), "nnz": int(csr.nnz),
                                            "h2d_bytes": int(csr.nnz * 8 + (N + 1) * 8)}
        del qs, csr

    cpu = None
    if rank == 0 and not args.no_cpu:
        if full_affinity is not None:   # the CPU baseline runs on all host cores, not on this GPU's NUMA node only
            os.sched_setaffinity(0, full_affinity)
        cpu = cpu_baseline(w, args.cpu_n_map, args.cpu_n_knn, ref_host=(Zref, comp, C, Nr))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "name": args.workload, "cells_per_gpu": N, "ref_cells": M, "genes": G,
                       "pcs": d, "clusters": K, "k": k, "batch_levels": levels,
                       "l2": "inputs larger than L2 (X is %.1f GB per step)" % (N * G * 4 / 1e9),
                       "knn_method": args.knn_method, "numa_node": numa_node},
            "stage_ms": stage_ms, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if extras is not None:
            line["extras"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
