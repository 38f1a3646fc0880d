"""Repeated cold map_embedding of a config3-shaped sample against the oracle: which rows / columns differ?"""
import copy, os, sys, warnings, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from symphonypy_b200 import synth, tl
from oracle import symphony_oracle as so
logging.getLogger("symphonypy").setLevel(logging.ERROR)
N, M, G, d, K = [int(x) for x in os.environ.get("SHAPE", "10000,20000,2000,50,200").split(",")]
query, ref = synth.small_case(N=N, M=M, G=G, d=d, K=K, L=12, levels=(8,), seed=3)
qo = copy.deepcopy(query)
INPUT = os.environ.get("INPUT", "pageable")
if INPUT == "pinned":
    t = torch.empty(query.X.shape, dtype=torch.float32, pin_memory=True); t.copy_(torch.from_numpy(np.asarray(query.X, dtype=np.float32)))
    query.X = t.numpy()
elif INPUT == "device":
    query.X = torch.from_numpy(np.asarray(query.X, dtype=np.float32)).cuda()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    so.map_embedding(qo, ref, key="batch0")
want = qo.obsm["X_pca_reference"]
for trial in range(int(os.environ.get("TRIALS", 6))):
    q = copy.deepcopy(query) if INPUT == 'pageable' else copy.copy(query)
    if trial % 2 == 0:
        tl.clear_caches()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tl.map_embedding(q, ref, key="batch0")
    got = q.obsm["X_pca_reference"]
    err = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    bad = np.where(err > 1e-4)[0]
    rerr = np.abs(q.obsm["X_pca_harmony_symphony_R"] - qo.obsm["X_pca_harmony_symphony_R"]).max()
    print(f"trial {trial} cold={trial % 2 == 0}: proj max {err.max():.3e}  bad rows {len(bad)} {bad[:12].tolist()}  R_abs {rerr:.2e}")
    if len(bad):
        r = bad[0]
        cols = np.where(np.abs(got[r] - want[r]) > 1e-4 * np.abs(want[r]).max())[0]
        print("   row", r, "bad columns", cols[:20].tolist(), "got", got[r, cols[:4]], "want", want[r, cols[:4]])
