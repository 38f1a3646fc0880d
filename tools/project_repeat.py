"""The projection kernel alone, launched repeatedly on the same device-resident input: every launch must return the
same bits (a data race inside the kernel shows up as rows that change between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from symphonypy_b200 import engine, synth
dev = torch.device("cuda", 0)
N, G, REPS = int(os.environ.get("N", 200_000)), 2000, int(os.environ.get("REPS", 200))
for d in [int(x) for x in os.environ.get("DS", "50,30,100").split(",")]:
    model = synth.make_model(G, d, 12, 0, dev)
    X, codes, _ = synth.make_query(model, N, [1], 7)
    pad = (-d) % 4
    ld = engine.Loadings(mu=model.mean, inv_sd=1.0 / model.std, U=torch.nn.functional.pad(model.U, (0, pad)), d=d)
    Z0 = engine.project_dense(X, ld).clone()
    want = ((X.double() - model.mean.double()) / model.std.double()).clamp(-10, 10) @ model.U.double()
    rel0 = ((Z0.double() - want).norm(dim=1) / want.norm(dim=1)).max().item()
    bad_launches, rows = 0, set()
    junk = torch.empty(64 << 20, device=dev)
    for i in range(REPS):
        if i % 3 == 0:
            junk.normal_()                       # perturb timing / caches between launches
        Z = engine.project_dense(X, ld)
        diff = (Z != Z0).any(1)
        nb = int(diff.sum())
        if nb:
            bad_launches += 1
            rows.update((torch.nonzero(diff).flatten() % 128).tolist()[:16])
    print(f"d={d} TMA={os.environ.get('SYM_PROJECT_TMA', '1')}: first launch rel err {rel0:.2e}; {bad_launches} of {REPS} launches differ; rows-in-tile {sorted(rows)[:24]}")
