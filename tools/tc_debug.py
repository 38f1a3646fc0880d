"""GPU-side diagnosis of the tcgen05 score tile (prints got/want so layout mistakes can be read off)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symphonypy_b200 import _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
import itertools
for d, ts in itertools.product((20, 30, 50, 7), (0, 1)):
    q = (rng.normal(size=(100, d)) * 3).astype(np.float32)
    r = (rng.normal(size=(128, d)) * 3).astype(np.float32)
    Q, R = torch.from_numpy(q).to(dev), torch.from_numpy(r).to(dev)
    out = torch.full((128, 128), -7.0, device=dev)
    scratch = torch.zeros(3 * 256 * 256 + 1024, dtype=torch.uint8, device=dev)
    rc = lib.sym_debug_tc_scores(ctypes.c_void_p(Q.data_ptr()), 100, ctypes.c_void_p(R.data_ptr()), 128, d,
                                 ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(scratch.data_ptr()), None, ts)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = (r.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * q.astype(np.float64) @ r.astype(np.float64).T
    scale = 2 * np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(r, axis=1)[None, :] + (r ** 2).sum(1)[None, :]
    err = np.abs(got[:100] - want) / scale
    print(f"d={d} ts={ts} rc={rc} max_rel_err={err.max():.3e} median={np.median(err):.3e}")
    if err.max() > 1e-4:
        np.set_printoptions(precision=2, linewidth=200, suppress=True)
        print("got[0:4,0:8]\n", got[0:4, 0:8], "\nwant[0:4,0:8]\n", want[0:4, 0:8])
        print("got.T match?", np.abs(got[:100, :100].T - want[:100, :100]).max())
        # per-element hypothesis: which want entry is each got entry closest to (first row)
        for i in (0, 1, 9):
            j = [int(np.argmin(np.abs(want[i] - got[i, c]))) for c in range(12)]
            print(f"row {i}: got col c matches want col", j)
        norms_only = (r.astype(np.float64) ** 2).sum(1)
        print("norm-only match", np.abs(got[0] - norms_only).max())
