"""Multi-GPU parity check (run under torchrun): every rank maps its row shard; the all-reduced MoE statistics
must make the sharded result equal the oracle's result on the WHOLE query."""
import copy, logging, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle import symphony_oracle as so
from symphonypy_b200 import synth, tl

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tl.settings.device = torch.device("cuda", local)
logging.getLogger("symphonypy").setLevel(logging.ERROR); logging.getLogger("symphony_oracle").setLevel(logging.ERROR)
query, ref = synth.small_case(N=4000, M=6000, G=512, d=20, K=40, L=8, levels=(5, 3), seed=7, missing_frac=0.1)
n = len(query); lo, hi = rank * n // world, (rank + 1) * n // world
shard = query[np.arange(lo, hi), :]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    tl.map_embedding(shard, ref, key=["batch0", "batch1"], lamb=[1.0, 2.0])
    tl.transfer_labels_kNN(shard, ref, "cell_type", n_neighbors=10)
    full = copy.deepcopy(query)
    so.map_embedding(full, ref, key=["batch0", "batch1"], lamb=[1.0, 2.0])
    so.transfer_labels_knn(full, ref, "cell_type", n_neighbors=10)
a, b = shard.obsm["X_pca_harmony"].astype(np.float64), full.obsm["X_pca_harmony"][lo:hi]
err = float((np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max())
same = float((np.asarray(shard.obs["cell_type"]) == np.asarray(full.obs["cell_type"])[lo:hi]).mean())
res = torch.tensor([err, 1.0 - same], device="cuda", dtype=torch.float64)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dist_check world={world}: max row-relative error of the corrected embedding {res[0].item():.2e}, label mismatch {res[1].item():.4f}")
    assert res[0].item() < 1e-4 and res[1].item() <= 1e-3
dist.destroy_process_group()
