"""Why does (or doesn't) tile pruning skip tiles?  Coarse-cluster radii of a clustered synthetic reference and the
tiles scanned with pruning on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from symphonypy_b200 import engine

dev = torch.device("cuda", 0)
N, M, d, k, sep = int(os.environ.get("N", 9000)), int(os.environ.get("M", 40000)), 30, 10, float(os.environ.get("SEP", 6.0))
rng = np.random.default_rng(N + M)
cent = rng.normal(0, 1.0, size=(16, d)) * sep
ref = torch.from_numpy((cent[rng.integers(0, 16, M)] + rng.normal(size=(M, d))).astype(np.float32)).to(dev)
qry = torch.from_numpy((cent[rng.integers(0, 16, N)] + rng.normal(size=(N, d))).astype(np.float32)).to(dev)
index = engine.knn_fit(ref)
code = engine.nearest_centroid(ref, index.centroids).long()
C = index.centroids.shape[0]
rad, cnt = [], []
for c in range(C):
    x = ref[code == c]
    cnt.append(int(x.shape[0]))
    rad.append(float((x - x.mean(0)).norm(dim=1).max()) if x.shape[0] else 0.0)
print("coarse clusters:", C, "sizes", cnt)
print("radii", [round(r, 1) for r in rad])
for method in (engine.KNN_TC16, engine.KNN_TC):
    for pr in (False, True):   # first tier alone (no repair passes)
        engine.knn_search(index, qry, k, method, prune=pr)
        ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta.record(); _, _, u = engine.knn_search(index, qry, k, method, prune=pr); tb.record(); torch.cuda.synchronize()
        print(f"  method {method} prune={pr}: first tier {ta.elapsed_time(tb):.2f} ms, unsafe rows {int(u.sum())}")
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.knn_search_certified(index, qry, k, method)
    t0.record(); i0, d0, _ = engine.knn_search_certified(index, qry, k, method); t1.record(); torch.cuda.synchronize()
    ms0 = t0.elapsed_time(t1)
    engine.knn_search_certified(index, qry, k, method, prune=True)
    t0.record(); i1, d1, _ = engine.knn_search_certified(index, qry, k, method, prune=True); t1.record(); torch.cuda.synchronize()
    print(f"method {method}: full {ms0:.2f} ms, pruned {t0.elapsed_time(t1):.2f} ms, tiles {index.scan_stats.tolist()}, same {bool(torch.equal(i0, i1))}")
