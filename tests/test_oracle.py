"""The numpy oracle against (a) golden vectors produced by the unmodified reference and
(b) the reference itself, live, when /root/reference is present (build container only)."""
import copy
import glob
import logging
import os
import warnings

import numpy as np
import pytest

from oracle import golden_io, ref_loader, symphony_oracle as so
from symphonypy_b200 import synth

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("confidence_"))


def _run_oracle(query, ref, call):
    q = copy.deepcopy(query)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        so.map_embedding(q, ref, **copy.deepcopy(call["map_embedding"]))
        kw = dict(call["transfer"])
        so.transfer_labels_knn(q, ref, kw.pop("ref_labels"), **kw)
    return q


def test_golden_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    logging.getLogger("symphony_oracle").setLevel(logging.ERROR)
    query, ref, call, out = golden_io.unpack(path)
    q = _run_oracle(query, ref, call)
    # the oracle repeats the reference's float64 numpy arithmetic: agreement is to rounding noise of BLAS
    np.testing.assert_allclose(q.obsm["X_pca_reference"], out["X_pca_reference"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(q.obsm["X_pca_harmony"], out["X_pca_harmony"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(q.obsm["X_pca_harmony_symphony_R"], out["R"], rtol=0, atol=1e-12)
    for k, v in out.items():
        if k.startswith("label__"):
            assert (np.asarray(q.obs[k[7:]], dtype=str) == v).all()


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("case", [
    dict(kw=dict(N=400, M=900, G=256, d=20, K=20, L=6, levels=(1,), seed=21), key=None, lamb=None),
    dict(kw=dict(N=300, M=800, G=200, d=25, K=12, L=5, levels=(5, 3, 2), seed=22, missing_frac=0.2,
                 sparse=True, extra_query_genes=17), key=["batch0", "batch1", "batch2"], lamb=[1.0, 0.5, 3.0]),
    dict(kw=dict(N=150, M=400, G=128, d=20, K=8, L=4, levels=(2,), seed=23, dtype64=False), key="batch0",
         lamb=[0.7, 1.3]),
])
def test_oracle_matches_live_reference(case):
    tl, _ = ref_loader.load()
    logging.getLogger("symphonypy").setLevel(logging.ERROR)
    logging.getLogger("symphony_oracle").setLevel(logging.ERROR)
    query, ref = synth.small_case(second_label=True, **case["kw"])
    qa, qb = copy.deepcopy(query), copy.deepcopy(query)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tl.map_embedding(qa, ref, key=case["key"], lamb=copy.deepcopy(case["lamb"]))
        so.map_embedding(qb, ref, key=case["key"], lamb=copy.deepcopy(case["lamb"]))
        tl.transfer_labels_kNN(qa, ref, ["cell_type", "lineage"], n_neighbors=9)
        so.transfer_labels_knn(qb, ref, ["cell_type", "lineage"], n_neighbors=9)
    for key in ("X_pca_reference", "X_pca_harmony", "X_pca_harmony_symphony_R"):
        assert qa.obsm[key].dtype == qb.obsm[key].dtype
        np.testing.assert_array_equal(qa.obsm[key], qb.obsm[key])
    for c in ("cell_type", "lineage"):
        assert (qa.obs[c] == qb.obs[c]).all()


def test_vote_rule_matches_sklearn():
    """`knn_vote_from_indices` restates sklearn's predict / predict_proba (ties -> lowest class)."""
    from sklearn.neighbors import KNeighborsClassifier

    rng = np.random.default_rng(0)
    ref = rng.normal(size=(500, 20))
    q = rng.normal(size=(300, 20))
    y = rng.integers(0, 5, 500)
    knn = KNeighborsClassifier(n_neighbors=10).fit(ref, y)
    idx, _ = so.knn_indices(ref, q, 10)
    win, frac = so.knn_vote_from_indices(idx, y, 5)
    assert (win == knn.predict(q)).all()
    np.testing.assert_allclose(frac, knn.predict_proba(q).max(axis=1))


def test_assign_sign_flip_quirk():
    """Rows whose entries are all negative are flipped by the signed row max (`_utils.py:168`)."""
    X = np.array([[-1.0, -2.0, -3.0], [1.0, 2.0, 3.0]])
    Y = np.eye(3)
    R = so.assign_clusters(X, 0.1, Y, 3)
    np.testing.assert_allclose(R[:, 0], R[:, 1])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_per_cell_confidence_oracle_matches_live_reference():
    tl, _ = ref_loader.load()
    logging.getLogger("symphonypy").setLevel(logging.ERROR)
    query, ref = synth.small_case(N=300, M=900, G=200, d=20, K=12, L=5, levels=(3,), seed=31, keep_R=True)
    r1, r2 = copy.deepcopy(ref), copy.deepcopy(ref)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tl.map_embedding(query, r1, key="batch0")
        q2 = copy.deepcopy(query)
        tl.per_cell_confidence(query, r1)
        so.per_cell_confidence(q2, r2)
    np.testing.assert_allclose(np.asarray(q2.obs["symphony_per_cell_dist"]), np.asarray(query.obs["symphony_per_cell_dist"]),
                               rtol=1e-10)
    np.testing.assert_allclose(r2.uns["harmony"]["inv_cluster_covs"], r1.uns["harmony"]["inv_cluster_covs"], rtol=1e-9)


def test_per_cell_confidence_golden():
    """Fixture made by the unmodified reference (oracle/gen_golden.py)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "confidence_small.npz")
    z = np.load(path)
    from symphonypy_b200.anndata_lite import AnnDataLite
    import pandas as pd
    q = AnnDataLite(np.zeros((z["Xq"].shape[0], 1), np.float32),
                    obsm={"X_pca_reference": z["Xq"], "X_pca_harmony_symphony_R": z["Rq"]})
    r = AnnDataLite(np.zeros((z["Xref"].shape[0], 1), np.float32), obsm={"X_pca_harmony": z["Xref"]},
                    uns={"harmony": {"R": z["Rref"]}})
    so.per_cell_confidence(q, r)
    np.testing.assert_allclose(np.asarray(q.obs["symphony_per_cell_dist"]), z["dist"], rtol=1e-10)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_per_cluster_confidence_oracle_matches_live_reference():
    import pandas as pd
    from symphonypy_b200.anndata_lite import AnnDataLite
    tl, _ = ref_loader.load()
    logging.getLogger("symphonypy").setLevel(logging.ERROR)
    rng = np.random.default_rng(5)
    n, d, K = 900, 6, 9
    obs = pd.DataFrame({"cl": rng.integers(0, 5, n).astype(str), "b": rng.integers(0, 2, n).astype(str)},
                       index=[f"c{i}" for i in range(n)])
    obs.loc[obs.index[:4], "cl"] = "tiny"
    q = AnnDataLite(np.zeros((n, 1), np.float32), obs=obs, obsm={"X_pca_reference": rng.normal(size=(n, d)) + 2})
    r = AnnDataLite(np.zeros((2, 1), np.float32),
                    uns={"harmony": {"C": rng.normal(size=(K, d)) * 10, "Nr": rng.uniform(5, 20, K)}})
    for key, lamb in (("cl", 0), (["cl", "b"], 0.25)):
        q1, q2 = copy.deepcopy(q), copy.deepcopy(q)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tl.per_cluster_confidence(q1, r, key, lamb=lamb, obs=None)   # obs: KeyError under pandas >= 3 (tl.py:179)
            so.per_cluster_confidence(q2, r, key, lamb=lamb)
        a, b = q1.uns["symphony_per_cluster_dist"], q2.uns["symphony_per_cluster_dist"]
        assert list(a["cluster_labels"]) == list(b["cluster_labels"])
        np.testing.assert_allclose(np.asarray(a["dist"], dtype=np.float64), b["dist"], rtol=1e-12, equal_nan=True)
        assert np.isnan(b["dist"]).any()


def test_per_cluster_confidence_golden():
    """Fixture made by the unmodified reference (oracle/gen_golden.py::cluster_confidence_fixture)."""
    import pandas as pd
    from symphonypy_b200.anndata_lite import AnnDataLite
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "confidence_cluster_small.npz"))
    n = z["Xq"].shape[0]
    q = AnnDataLite(np.zeros((n, 1), np.float32),
                    obs=pd.DataFrame({"cell_type": z["cell_type"], "batch0": z["batch0"]}, index=pd.RangeIndex(n).astype(str)),
                    obsm={"X_pca_reference": z["Xq"]})
    r = AnnDataLite(np.zeros((2, 1), np.float32), uns={"harmony": {"C": z["C"], "Nr": z["Nr"]}})
    so.per_cluster_confidence(q, r, "cell_type")
    np.testing.assert_allclose(q.uns["symphony_per_cluster_dist"]["dist"], z["dist_a"], rtol=1e-10, equal_nan=True)
    so.per_cluster_confidence(q, r, ["cell_type", "batch0"], lamb=0.5, u=1)
    np.testing.assert_allclose(q.uns["symphony_per_cluster_dist"]["dist"], z["dist_b"], rtol=1e-10, equal_nan=True)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_soft_kmeans_summary_matches_live_reference():
    """The reference's fallback (`_utils.py:311-352`) is random in its centres; given those centres every stored
    field must be reproduced exactly."""
    _, utils = ref_loader.load()
    _, ref = synth.small_case(N=20, M=600, G=64, d=10, K=5, L=4, seed=3)
    ref.obsm["X_pca"] = np.asarray(ref.obsm["X_pca_harmony"], dtype=np.float64)
    del ref.uns["harmony"]
    utils._run_soft_kmeans(ref, ref_basis="X_pca", K=None, sigma=0.1)
    h = ref.uns["harmony"]
    assert h["K"] == 20 and h["C"].shape == (20, 10)
    mine = so.soft_kmeans_summary(ref.obsm["X_pca"], h["C"], 0.1)
    assert set(mine) == set(h)
    np.testing.assert_allclose(mine["R"], h["R"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(mine["Nr"], h["Nr"], rtol=1e-12)
    for k in ("K", "sigma", "ref_basis_loadings", "ref_basis_adjusted", "vars_use", "harmony_kwargs", "converged"):
        assert mine[k] == h[k]


def test_kmeans_plusplus_restatement_matches_sklearn():
    """The oracle's greedy k-means++ (explicit random draws) picks the rows scikit-learn's own `kmeans_plusplus`
    picks when it is fed the numbers sklearn's RandomState would draw: pins the restatement to sklearn 1.9.0
    (`symphonypy/_utils.py:323-325` calls it through KMeans)."""
    from sklearn.cluster import kmeans_plusplus

    rng = np.random.default_rng(5)
    for n, d, K, seed in ((600, 8, 12, 0), (2500, 20, 40, 3), (300, 5, 100, 11)):
        X = np.concatenate([rng.normal(c, 0.6, size=(n // 4, d)) for c in (-3, 0, 2, 5)]).astype(np.float32).astype(np.float64)
        _, want = kmeans_plusplus(X, K, random_state=seed)
        # the same stream of draws: choice(n) for the first centre, then uniform(size=trials) per centre
        rs = np.random.RandomState(seed)
        trials = 2 + int(np.log(K))
        first = rs.choice(X.shape[0], p=np.full(X.shape[0], 1.0 / X.shape[0]))
        rand = np.stack([rs.uniform(size=trials) for _ in range(K - 1)])
        got = so.kmeans_plusplus_ids(X, K, first, rand)
        assert np.array_equal(got, want)
