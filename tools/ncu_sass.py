"""Per-instruction view of an ncu capture: SASS with stall samples and the dominant stall reasons.

    python tools/ncu_sass.py gpurun_out/prof_<tag>.ncu-rep [min_samples_pct] > profiles/<tag>_sass.txt
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 6]
h = rows[0]
si, ci, ei, ti = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
stall = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[1:] if r[ci].isdigit()]
tot = sum(int(r[ci]) for r in body) or 1
agg = {}
print(f"# {rep}: {tot} warp-stall samples; columns: addr-offset  samples%  executed  avg-threads  SASS  | top stall reasons")
base = int(body[0][0], 16)
for r in body:
    c = int(r[ci])
    top = sorted(((int(r[i]), n[6:]) for i, n in stall if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    for v, n in top:
        agg[n] = agg.get(n, 0) + v
    if 100.0 * c / tot >= minpct:
        print(f"{int(r[0],16)-base:05x} {100.0*c/tot:6.2f} {r[ei]:>10} {r[ti]:>5}  {r[si].strip():<70} | " + " ".join(f"{n}:{v}" for v, n in top))
print("# stall reasons over all instructions (top-3 per instruction summed):", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
