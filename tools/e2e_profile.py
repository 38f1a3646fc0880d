"""Host-side profile of the public API on config2 (where does the end-to-end time go?)."""
import cProfile, pstats, sys, os, time, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from symphonypy_b200 import synth, tl, engine
from symphonypy_b200.anndata_lite import AnnDataLite
logging.getLogger("symphonypy").setLevel(logging.ERROR)
dev = torch.device("cuda", 0)
N, M, G, d, K, k = int(os.environ.get("N", 1_000_000)), 500_000, 2000, 30, 100, 10
model = synth.make_model(G, d, 20, 0, dev)
Zref, comp, C, Nr = synth.make_reference(model, M, K, 0)
X, codes, _ = synth.make_query(model, N, [1], 1000)
FMT = os.environ.get("FMT", "pinned")   # pinned | pageable | csr | csr_pinned
Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=(FMT == "pinned")); Xh.copy_(X); del X
Xq = Xh.numpy()
if FMT in ("csr", "csr_pinned"):
    import scipy.sparse as sp
    Xq = sp.csr_matrix(Xq)
    if FMT == "csr_pinned":
        dh = torch.empty(Xq.nnz, dtype=torch.float32, pin_memory=True); dh.copy_(torch.from_numpy(Xq.data))
        ih = torch.empty(Xq.nnz, dtype=torch.int32, pin_memory=True); ih.copy_(torch.from_numpy(Xq.indices))
        Xq = sp.csr_matrix((dh.numpy(), ih.numpy(), Xq.indptr), shape=Xq.shape)
q = AnnDataLite(Xq, obs=pd.DataFrame(index=pd.RangeIndex(N).astype(str)),
                var=pd.DataFrame(index=pd.Index([f"g{i}" for i in range(G)])), uns={"log1p": {}})
_, ref = synth.to_anndata(model, Zref, comp, C, Nr, Xh[:8].to(dev), codes[:, :8], K, seed=0)
def step():
    tl.map_embedding(q, ref, key=None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    tl.transfer_labels_kNN(q, ref, "cell_type", n_neighbors=k)
    torch.cuda.synchronize()
    return t1
step()
t0 = time.perf_counter(); t1 = step(); t2 = time.perf_counter()
print(f"map_embedding {1e3*(t1-t0):.1f} ms, transfer {1e3*(t2-t1):.1f} ms")
if os.environ.get("COLD") == "1":   # what the `cold` variant of bench.py times: every cache and staging buffer dropped first
    tl.clear_caches(); torch.cuda.synchronize()
    t0 = time.perf_counter(); t1 = step(); t2 = time.perf_counter()
    print(f"cold: map_embedding {1e3*(t1-t0):.1f} ms, transfer {1e3*(t2-t1):.1f} ms")
    tl.clear_caches(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(int(os.environ.get("TOP", 28)))
