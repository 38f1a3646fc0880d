"""Wall-clock timeline of one warm end-to-end step (config2, CSR arrays in page-locked memory): entry / exit of the
host functions on the path, no extra synchronisation.  Usage: python tools/e2e_timeline.py"""
import os, sys, time, logging, functools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch, scipy.sparse as sp
from symphonypy_b200 import synth, tl, engine, pipeline, _design
from symphonypy_b200.anndata_lite import AnnDataLite
logging.getLogger("symphonypy").setLevel(logging.ERROR)
dev = torch.device("cuda", 0)
N, M, G, d, K, k = int(os.environ.get("N", 1_000_000)), 500_000, 2000, 30, 100, 10
model = synth.make_model(G, d, 20, 0, dev)
Zref, comp, C, Nr = synth.make_reference(model, M, K, 0)
X, codes, _ = synth.make_query(model, N, [1], 1000)
Xq = sp.csr_matrix(X.cpu().numpy())
dh = torch.empty(Xq.nnz, dtype=torch.float32, pin_memory=True); dh.copy_(torch.from_numpy(Xq.data))
ih = torch.empty(Xq.nnz, dtype=torch.int32, pin_memory=True); ih.copy_(torch.from_numpy(Xq.indices))
Xq = sp.csr_matrix((dh.numpy(), ih.numpy(), Xq.indptr), shape=Xq.shape)
q = AnnDataLite(Xq, obs=pd.DataFrame(index=pd.RangeIndex(N).astype(str)),
                var=pd.DataFrame(index=pd.Index([f"g{i}" for i in range(G)])), uns={"log1p": {}})
_, ref = synth.to_anndata(model, Zref, comp, C, Nr, X[:8], codes[:, :8], K, seed=0)
del X
LOG, T0 = [], [0.0]
def wrap(mod, name):
    f = getattr(mod, name)
    @functools.wraps(f)
    def g(*a, **kw):
        t = time.perf_counter()
        try:
            return f(*a, **kw)
        finally:
            LOG.append((name, 1e3 * (t - T0[0]), 1e3 * (time.perf_counter() - T0[0])))
    setattr(mod, name, g)
for mod, names in ((tl, ["_project", "_encode_labels", "_decode_labels", "_to_host", "_to_host_async", "_device_copy",
                         "content_digest", "_map_embedding_once"]),
                   (pipeline, ["soft_assign_and_correct", "check_solve"]),
                   (_design, ["batch_codes"]),
                   (engine, ["knn_search_certified", "knn_search", "vote", "nearest_centroid"])):
    for n in names:
        if hasattr(mod, n):
            wrap(mod, n)
def step():
    tl.map_embedding(q, ref, key=None)
    t = 1e3 * (time.perf_counter() - T0[0]); LOG.append(("map_embedding returned", t, t))
    tl.transfer_labels_kNN(q, ref, "cell_type", n_neighbors=k)
    torch.cuda.synchronize()
for _ in range(3):
    step()
for rep in range(2):
    LOG.clear(); torch.cuda.synchronize(); T0[0] = time.perf_counter(); step()
    total = 1e3 * (time.perf_counter() - T0[0])
    print(f"--- step {rep}: {total:.1f} ms")
    for name, a, b in sorted(LOG, key=lambda r: r[1]):
        print(f"{a:8.2f} -> {b:8.2f}  ({b - a:6.2f})  {name}")
