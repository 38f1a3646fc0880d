"""Repeated runs of one scan method on one index: flagged rows and neighbour lists must not change."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from symphonypy_b200 import engine
dev = torch.device("cuda", 0)
N, M, d, k = [int(x) for x in os.environ.get("SHAPE", "20000,50000,30,10").split(",")]
rng = np.random.default_rng(M + k)
centres = rng.normal(size=(8, d)) * 4
ref = (centres[rng.integers(0, 8, M)] + rng.normal(size=(M, d))).astype(np.float32)
q = (centres[rng.integers(0, 8, N)] + rng.normal(size=(N, d))).astype(np.float32)
q[:3] *= 1e4
index = engine.knn_fit(torch.from_numpy(ref).to(dev))
Q = torch.from_numpy(q).to(dev)
for m in (3, 2):
    prev = None
    for rep in range(6):
        if rep == 3:
            index.workspace.clear()                      # fresh (uninitialised) workspace
            junk = torch.full((200_000_000,), 0x7f, dtype=torch.uint8, device=dev); del junk
        idx, dist2, unsafe = engine.knn_search(index, Q, k, m)
        torch.cuda.synchronize()
        cur = (unsafe.clone(), idx.clone())
        msg = f"method {m} rep {rep}: flagged {int(unsafe.sum())}"
        if prev is not None:
            du = int((prev[0] != cur[0]).sum()); di = int((prev[1] != cur[1]).any(1).sum())
            msg += f"  rows with different flag {du}, different idx {di}"
            if du:
                rows = torch.nonzero(prev[0] != cur[0]).flatten()[:8].tolist()
                msg += f" rows {rows}"
        print(msg)
        prev = cur

