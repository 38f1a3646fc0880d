"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden
Each fixture holds the inputs, the call arguments and the reference's outputs
(`obsm[X_pca_reference]`, `obsm[X_pca_harmony]`, `obsm[..._symphony_R]`, and the
labels `transfer_labels_kNN` wrote).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import copy
import logging
import os
import warnings

import numpy as np

from oracle import golden_io, ref_loader
from symphonypy_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (synthetic-case kwargs, map_embedding kwargs, transfer kwargs)
CASES = {
    # all genes present, dense float32 query, no batch key (Phi = two rows of ones)
    "dense_f32_nokey": (
        dict(N=256, M=600, G=300, d=20, K=25, L=6, levels=(1,), seed=3, dtype64=False),
        dict(key=None),
        dict(ref_labels="cell_type", n_neighbors=10)),
    # CSR query, 10 % genes missing, zero-std genes, interleaved foreign genes, two batch keys
    "csr_missing_twokeys": (
        dict(N=300, M=700, G=320, d=24, K=30, L=7, levels=(3, 12), seed=5, missing_frac=0.10,
             zero_std_frac=0.02, extra_ref_genes=40, extra_query_genes=30, sparse=True, second_label=True),
        dict(key=["batch0", "batch1"], lamb=[0.5, 2.0]),
        dict(ref_labels=["cell_type", "lineage"], n_neighbors=7)),
    # one key given as str, scalar lamb, per-cluster sigma array, default k=5
    "dense_onekey_sigma_array": (
        dict(N=200, M=500, G=256, d=30, K=16, L=5, levels=(4,), seed=9, missing_frac=0.05),
        dict(key="batch0", lamb=0.25, sigma=list(np.linspace(0.08, 0.2, 16))),
        dict(ref_labels=["cell_type"], query_labels=["predicted"])),
}


def main():
    assert ref_loader.available(), "needs /root/reference"
    tl, _ = ref_loader.load()
    logging.getLogger("symphonypy").setLevel(logging.ERROR)
    os.makedirs(OUT, exist_ok=True)
    for name, (case, me_kw, knn_kw) in CASES.items():
        query, ref = synth.small_case(**case)
        q = copy.deepcopy(query)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tl.map_embedding(q, ref, **copy.deepcopy(me_kw))
            kw = dict(knn_kw)
            ref_labels = kw.pop("ref_labels")
            tl.transfer_labels_kNN(q, ref, ref_labels, **kw)
        qcols = kw.get("query_labels", ref_labels)
        qcols = [qcols] if isinstance(qcols, str) else qcols
        outputs = {"X_pca_reference": q.obsm["X_pca_reference"], "X_pca_harmony": q.obsm["X_pca_harmony"],
                   "R": q.obsm["X_pca_harmony_symphony_R"]}
        for c in qcols:
            outputs[f"label__{c}"] = np.asarray(q.obs[c], dtype=str)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **golden_io.pack(query, ref, {"map_embedding": me_kw, "transfer": knn_kw}, outputs))
        print(name, os.path.getsize(path) // 1024, "KiB")


def confidence_fixture():
    """per_cell_confidence (tl.py:33-111): inputs at the function's boundary and the reference's output."""
    tl, _ = ref_loader.load()
    query, ref = synth.small_case(N=200, M=500, G=160, d=20, K=10, L=5, levels=(2,), seed=41, keep_R=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tl.map_embedding(query, ref, key="batch0")
        tl.per_cell_confidence(query, ref)
    path = os.path.join(OUT, "confidence_small.npz")
    np.savez_compressed(path, Xq=query.obsm["X_pca_reference"], Rq=query.obsm["X_pca_harmony_symphony_R"],
                        Xref=ref.obsm["X_pca_harmony"], Rref=ref.uns["harmony"]["R"],
                        dist=np.asarray(query.obs["symphony_per_cell_dist"]))
    print("confidence_small", os.path.getsize(path) // 1024, "KiB")


def cluster_confidence_fixture():
    """per_cluster_confidence (tl.py:114-179) on clusters = transferred labels; `obs=None` because the
    reference's per-cell broadcast (`tl.py:179`) raises KeyError under the pandas of this image."""
    tl, _ = ref_loader.load()
    query, ref = synth.small_case(N=400, M=700, G=160, d=12, K=10, L=4, levels=(2,), seed=43)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tl.map_embedding(query, ref, key="batch0")
        tl.transfer_labels_kNN(query, ref, "cell_type", n_neighbors=5)
        query.obs.loc[query.obs.index[:7], "cell_type"] = "rare"      # a cluster below u*d cells
        tl.per_cluster_confidence(query, ref, "cell_type", obs=None)
        a = dict(query.uns["symphony_per_cluster_dist"])
        tl.per_cluster_confidence(query, ref, ["cell_type", "batch0"], lamb=0.5, u=1, obs=None)
        b = dict(query.uns["symphony_per_cluster_dist"])
    path = os.path.join(OUT, "confidence_cluster_small.npz")
    np.savez_compressed(path, Xq=query.obsm["X_pca_reference"], C=ref.uns["harmony"]["C"], Nr=ref.uns["harmony"]["Nr"],
                        cell_type=np.asarray(query.obs["cell_type"], dtype=str),
                        batch0=np.asarray(query.obs["batch0"], dtype=str),
                        dist_a=np.asarray(a["dist"], dtype=np.float64),
                        labels_a=np.asarray(a["cluster_labels"], dtype=str),
                        dist_b=np.asarray(b["dist"], dtype=np.float64),
                        labels_b=np.asarray(["|".join(t) for t in b["cluster_labels"]], dtype=str))
    print("confidence_cluster_small", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
    confidence_fixture()
    cluster_confidence_fixture()
