"""Cycle accounting of the tcgen05 scan (library built with SYM_NVCC_EXTRA=-DTC_PROFILE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from symphonypy_b200 import engine, _lib

dev = torch.device("cuda", 0)
N, M, d, k = int(os.environ.get("N", 1_000_000)), int(os.environ.get("M", 500_000)), int(os.environ.get("D", 30)), int(os.environ.get("K", 10))
g = torch.Generator(device=dev); g.manual_seed(0)
ref = torch.randn((M, d), generator=g, device=dev)
q = torch.randn((N, d), generator=g, device=dev)
index = engine.knn_fit(ref)
lib = _lib.load()
out = (ctypes.c_ulonglong * 16)()
has_prof = hasattr(lib, "sym_debug_tc_profile")   # only in -DTC_PROFILE builds
for rep in range(2):
    if has_prof:
        lib.sym_debug_tc_profile(out)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    engine.knn_search(index, q, k, int(os.environ.get("METHOD", 3)))
    t1.record(); torch.cuda.synchronize()
    if has_prof:
        lib.sym_debug_tc_profile(out)
if not has_prof:
    print(f"knn_search {t0.elapsed_time(t1):.1f} ms (no cycle counters in this build)")
    sys.exit(0)
v = list(out)
wt = max(v[4], 1)
print(f"knn_search {t0.elapsed_time(t1):.1f} ms; selection warp-tiles {v[4]}")
print(f"  per warp-tile cycles: wait_full {v[0]/wt:.0f}  ld {v[1]/wt:.0f}  examine {v[2]/wt:.0f} (mins + vote {v[7]/wt:.0f})  drain {v[3]/wt:.0f}  total {(v[0]+v[1]+v[2]+v[3])/wt:.0f}")
print(f"  warp-tiles with a hit {v[5]/wt:.3f}; lanes with a hit per warp-tile {v[6]/wt:.3f}; per 1000 warp-tiles: periodic drains {1000*v[14]/wt:.1f}, early drains {1000*v[15]/wt:.2f}")
it = max(v[12], 1)
print(f"  issuer per tile cycles: wait ref tile {v[8]/it:.0f}  wait free acc {v[9]/it:.0f}  wait turn {v[10]/it:.0f}  issue {v[11]/it:.0f}")
