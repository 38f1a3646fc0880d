"""numpy / scikit-learn restatement of the reference's query-mapping path.

TEST INFRASTRUCTURE (see `oracle/__init__.py`): the checker for the CUDA path and
the timed CPU baseline ("port") of `bench.py`.  Never imported by the product.

Each function names the reference lines it follows.  Array arithmetic is kept in
the same order and dtype as the reference so results agree to the last bit where
numpy is deterministic, and so that timing this port is representative of timing
the reference (the boolean-masked in-place scaling of `_utils.py:298-299` — the
reference's slowest step — is kept as such).

The kNN arithmetic of `tl.transfer_labels_kNN` is not in the reference tree: it
is scikit-learn's `KNeighborsClassifier` (unpinned in `setup.cfg:19-22`;
scikit-learn 1.9.0 in this image, brute force `EuclideanArgKmin64`,
SURVEY.md App. C).  `transfer_labels_knn` calls that estimator exactly as
`tl.py:425-436` does; `knn_vote_from_indices` restates its vote rule
(votes/k, argmax with ties to the lowest sorted class).
"""
from __future__ import annotations

import logging
import warnings

import numpy as np
import pandas as pd
from scipy.sparse import issparse

logger = logging.getLogger("symphony_oracle")

CLIP = 10.0  # `_utils.py:253`: max_value default, never overridden by `tl.py:350-356`


# --------------------------------------------------------------------------- project
def reference_genes(adata_ref, use_genes_column="highly_variable"):
    """`_utils.py:263-274`: gene list (reference order) and their stds."""
    if use_genes_column is None:
        return adata_ref.var_names, np.array(adata_ref.var["std"]), None
    hv = adata_ref.var[use_genes_column]
    return adata_ref.var_names[hv], np.array(adata_ref.var["std"][hv]), hv


def map_query_to_ref(adata_ref, adata_query, loadings_key="PCs", use_genes_column="highly_variable",
                     max_value=CLIP):
    """`_utils.py:248-308` (+ `_adjust_for_missing_genes`, `:217-245`).  Returns Z [N, d]."""
    assert "mean" in adata_ref.var, "Gene expression means are expected to be saved in adata_ref.var"
    assert "std" in adata_ref.var, "Gene expression stds are expected to be saved in adata_ref.var"
    genes, sd, hv = reference_genes(adata_ref, use_genes_column)
    present = genes.isin(adata_query.var_names) & (sd != 0)  # `:276`

    if present.all():  # `:283-285`: query dtype is kept
        block = adata_query[:, genes].X
        t = block.toarray() if issparse(block) else block.copy()
    else:  # `:233-243`: float64 zero buffer, present columns filled in
        logger.warning("%i out of %i genes from the reference are missing in the query dataset or "
                       "have zero std in the reference, their expressions in the query will be set to zero",
                       int((~present).sum()), genes.shape[0])
        t = np.zeros((adata_query.shape[0], genes.shape[0]))
        block = adata_query[:, genes[present]].X
        t[:, present] = block.toarray() if issparse(block) else block.copy()

    mu_all = adata_ref.var["mean"] if hv is None else adata_ref.var["mean"][hv]
    mu = np.array(mu_all[present])  # `:287-295`
    sd = sd[present]  # `:296`
    t[:, present] -= mu[None, :]  # `:298`
    t[:, present] /= sd[None, :]  # `:299`
    if max_value is not None:
        t = np.clip(t, -max_value, max_value)  # `:301-302`
    loadings = adata_ref.varm[loadings_key][adata_ref.var_names.isin(genes)]  # `:307`
    return np.array(t @ loadings)


# --------------------------------------------------------------------------- assign
def assign_clusters(X, sigma, Y, K):
    """`_utils.py:154-178`.  Returns R [K, N]."""
    if isinstance(sigma, (int, float)):
        sigma = np.array([sigma], dtype=np.float32)
    else:
        sigma = np.array(sigma, dtype=np.float32)
        assert len(sigma) == K, \
            "sigma paramater must be either a single float or an array of length equal to number of clusters"
    xc = X / X.max(axis=1, keepdims=True)  # signed row max (`:168`)
    xc /= np.linalg.norm(xc, ord=2, axis=1, keepdims=True)
    R = -2 * (1 - Y @ xc.T) / sigma[..., None]
    R -= R.max(axis=0)
    R = np.exp(R)
    R /= R.sum(axis=0, keepdims=True)
    return R


# --------------------------------------------------------------------------- design
def build_design(adata_query, key, lamb):
    """`tl.py:368-392`: Phi_ [(B+1), N] and the ridge matrix diag(0, lamb...)."""
    if key is None:
        batch = pd.DataFrame({"batch": ["1"] * len(adata_query)})
    elif isinstance(key, str):
        batch = pd.DataFrame(adata_query.obs[key]).astype(str)
    else:
        batch = adata_query.obs[key].astype(str)
    phi = pd.get_dummies(batch).to_numpy().T
    phi_ = np.concatenate([np.ones((1, phi.shape[1])), phi], axis=0)
    levels = batch.describe().loc["unique"].to_numpy().astype(int)
    if lamb is None:
        lamb = np.repeat([1] * len(levels), levels)
    elif isinstance(lamb, (float, int)):
        lamb = np.repeat([lamb] * len(levels), levels)
    elif len(lamb) == len(levels):
        lamb = np.repeat([lamb], levels)
    else:
        assert len(lamb) == np.sum(levels), "each batch variable must have a lambda"
    return phi_, np.diag(np.insert(lamb, 0, 0))


# --------------------------------------------------------------------------- MoE
def correct_query(X, phi_, R, Nr, C, lamb):
    """`_utils.py:181-214`.  Returns X_corr [N, d]."""
    out = X.copy().T
    lamb[0, 0] = 0
    for k in range(R.shape[0]):
        pr = np.multiply(phi_, R[k, :])  # `:196`
        gram = pr @ phi_.T  # `:199`
        gram[0, 0] += Nr[k]
        cross = pr @ X  # `:204`
        cross[0, :] += C[k]
        W = np.linalg.inv(gram + lamb) @ cross  # `:208`
        W[0, :] = 0
        out -= W.T @ pr  # `:212`
    return out.T


def moe_weights(X, phi_, R, Nr, C, lamb):
    """The per-cluster ridge weights W [K, B+1, d] of `_utils.py:196-209` (exposed for
    kernel-level tests; the reference never returns them)."""
    lamb = lamb.copy()
    lamb[0, 0] = 0
    W = np.zeros((R.shape[0], phi_.shape[0], X.shape[1]))
    for k in range(R.shape[0]):
        pr = phi_ * R[k, :]
        gram = pr @ phi_.T
        gram[0, 0] += Nr[k]
        cross = pr @ X
        cross[0, :] += C[k]
        W[k] = np.linalg.inv(gram + lamb) @ cross
        W[k, 0, :] = 0
    return W


# --------------------------------------------------------------------------- map_embedding
def map_embedding(adata_query, adata_ref, key=None, lamb=None, sigma=0.1,
                  use_genes_column="highly_variable", transferred_adjusted_basis="X_pca_harmony",
                  transferred_primary_basis="X_pca_reference"):
    """`tl.py:271-397` without the out-of-scope no-Harmony fallback (`:320-335`)."""
    assert "mean" in adata_ref.var, "Gene expression means are expected to be saved in adata_ref.var"
    assert "std" in adata_ref.var, "Gene expression stds are expected to be saved in adata_ref.var"
    assert "harmony" in adata_ref.uns, "oracle covers the Harmony-reference path only"
    if use_genes_column is not None:
        assert use_genes_column in adata_ref.var, \
            f"Column `{use_genes_column}` not found in adata_ref.var. Set `use_genes_column` parameter properly"
    if "log1p" not in adata_query.uns:
        warnings.warn("Gene expressions in adata_query should be log1p-transformed")
    h = adata_ref.uns["harmony"]
    Z = map_query_to_ref(adata_ref, adata_query, h["ref_basis_loadings"], use_genes_column)
    adata_query.obsm[transferred_primary_basis] = Z
    C = h["C"]
    Y = C / np.linalg.norm(C, ord=2, axis=1, keepdims=True)  # `:361-362`
    R = assign_clusters(Z, sigma, Y, h["K"])
    phi_, lam = build_design(adata_query, key, lamb)
    adata_query.obsm[transferred_adjusted_basis] = correct_query(Z, phi_, R, h["Nr"], h["C"], lam)
    adata_query.obsm[f"{transferred_adjusted_basis}_symphony_R"] = R.T


# --------------------------------------------------------------------------- per-cell confidence (next row)
def cluster_covs(X_ref, R, K):
    """`_utils.py:468-496`: inverse R-weighted covariance [K, d, d] and weighted mean [K, d, 1] of every
    reference cluster.  Same arithmetic; the [K, d, N_ref] temporaries of the reference are built per cluster
    so that a 500k-cell reference does not need 12 GB."""
    R = np.asarray(R)
    X_ref = np.asarray(X_ref)
    v1 = R.sum(axis=1, keepdims=True)
    v2 = (R ** 2).sum(axis=1, keepdims=True)
    d = X_ref.shape[1]
    mean = np.empty((K, d, 1))
    covs = np.empty((K, d, d))
    for k in range(K):
        mean[k, :, 0] = (R[k][None, :] * X_ref.T).sum(axis=1) / v1[k]
        centered = X_ref.T - mean[k]                       # [d, N_ref]
        covs[k] = (R[k][None, :] * centered) @ centered.T  # einsum("pq,rq->pr")
    covs = covs * v1[..., None] / (v1 ** 2 - v2)[..., None]
    return np.linalg.inv(covs), mean


def per_cell_confidence(adata_query, adata_ref, ref_basis_adjusted="X_pca_harmony", query_basis_adjusted="X_pca_harmony",
                        transferred_primary_basis="X_pca_reference", obs="symphony_per_cell_dist"):
    """`tl.py:33-111`: R-weighted Mahalanobis distance of every query cell to the reference clusters."""
    assert "harmony" in adata_ref.uns, \
        "Harmony object not found in adata_ref.uns['harmony']. First, run standard symphony query mapping."
    h = adata_ref.uns["harmony"]
    R = h["R"]
    K = R.shape[0]
    if ("inv_cluster_covs" not in h) or ("cluster_centers" not in h):
        h["inv_cluster_covs"], h["cluster_centers"] = cluster_covs(adata_ref.obsm[ref_basis_adjusted], R, K)
    inv_covs, centers = h["inv_cluster_covs"], h["cluster_centers"]
    Xq = adata_query.obsm[transferred_primary_basis]
    Rq = adata_query.obsm[f"{query_basis_adjusted}_symphony_R"].T
    out = np.zeros(Xq.shape[0])
    for k in range(K):  # the reference materialises [K, Nq, d]; same numbers cluster by cluster
        c = Xq - centers[k, :, 0][None, :]
        out += Rq[k] * np.sum((c @ inv_covs[k]) * c, axis=1) ** 0.5
    adata_query.obs[obs] = out


# --------------------------------------------------------------------------- per-cluster confidence (next row)
def cluster_maha_dist(query_coords, reference_cluster_centroids, reference_cluster_centroids_norm, u, lamb):
    """`_utils.py:422-465` for one query cluster given its coordinates [Nq, d]; None when Nq < u*d."""
    query_coords = np.asarray(query_coords)
    d = reference_cluster_centroids.shape[1]
    centroid = query_coords.mean(axis=0)
    centered = query_coords - centroid
    with np.errstate(divide="ignore", invalid="ignore"):
        cov = centered.T @ centered / (centered.shape[0] - 1)
    closest = reference_cluster_centroids[np.argmax(reference_cluster_centroids_norm @ centroid)]
    if query_coords.shape[0] < u * d:
        return None
    cov = cov + np.diag([lamb] * d)
    dif = closest - centroid
    return float((dif @ np.linalg.inv(cov) @ dif) ** 0.5)


def per_cluster_confidence(adata_query, adata_ref, cluster_key, u=2, lamb=0,
                           transferred_primary_basis="X_pca_reference", obs="symphony_per_cluster_dist",
                           uns="symphony_per_cluster_dist"):
    """`tl.py:114-179`.  The per-cell column is the broadcast the reference intends at `tl.py:179` (its
    `dists[x[0]]` raises KeyError under pandas >= 3, so the pinned part is `uns`)."""
    assert "harmony" in adata_ref.uns, \
        "Harmony object not found in adata_ref.uns['harmony']. First, run standard symphony query mapping."
    h = adata_ref.uns["harmony"]
    cent = h["C"] / h["Nr"][..., np.newaxis]
    cent_norm = cent / np.linalg.norm(cent, ord=2, axis=1, keepdims=True)
    X = np.asarray(adata_query.obsm[transferred_primary_basis])
    grouped = adata_query.obs.groupby(cluster_key)
    pos = grouped.indices  # label -> row positions
    labels = grouped.size().index
    dists = [cluster_maha_dist(X[pos[lab]], cent, cent_norm, u, lamb) for lab in labels]
    dist = np.array([np.nan if v is None else v for v in dists], dtype=np.float64)
    if uns is not None:
        adata_query.uns[uns] = {"key": cluster_key, "dist": dist, "cluster_labels": labels.to_numpy()}
    if obs is not None:
        code = grouped.ngroup().to_numpy()
        adata_query.obs[obs] = np.where(code >= 0, dist[np.clip(code, 0, len(dist) - 1)], np.nan)


# --------------------------------------------------------------------------- no-Harmony fallback (next row)
def soft_kmeans_summary(Z, C, sigma=0.1, ref_basis="X_pca", ref_basis_loadings="PCs"):
    """`_utils.py:327-352`: everything `_run_soft_kmeans` stores once the k-means centres C are known
    (the KMeans call itself, `_utils.py:323-325`, is unseeded sklearn and not reproducible)."""
    Z = np.asarray(Z)
    K = C.shape[0]
    Z_cos = Z.copy()
    Z_cos /= Z_cos.max(axis=1, keepdims=True)
    Z_cos /= np.linalg.norm(Z_cos, ord=2, axis=1, keepdims=True)
    Y = C / np.linalg.norm(C, ord=2, axis=1, keepdims=True)
    R = assign_clusters(Z_cos, sigma, Y, K)
    return {"Nr": R.sum(axis=1), "C": C, "K": K, "sigma": None, "ref_basis_loadings": ref_basis_loadings,
            "ref_basis_adjusted": ref_basis, "vars_use": None, "harmony_kwargs": {}, "converged": True, "R": R}


def kmeans_plusplus_ids(X, K, first, rand):
    """Greedy k-means++ as scikit-learn 1.9.0 runs it inside `KMeans(init="k-means++")` (`symphonypy/_utils.py:323-325`;
    sklearn/cluster/_kmeans.py, `_kmeans_plusplus`, unit sample weights), with the random numbers made explicit:
    `first` = row of the first centre, `rand[c-1]` = the `n_local_trials` uniform draws for centre c.
    Returns the chosen rows [K].  Pinned to sklearn in tests/test_oracle.py by feeding it sklearn's own draws."""
    from sklearn.metrics.pairwise import euclidean_distances

    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    x2 = np.einsum("ij,ij->i", X, X)
    ids = np.empty(K, dtype=np.int64)
    ids[0] = first
    closest = euclidean_distances(X[first][None, :], X, Y_norm_squared=x2, squared=True)[0]
    pot = closest.sum()
    for c in range(1, K):
        rand_vals = np.asarray(rand[c - 1], dtype=np.float64) * pot
        cand = np.searchsorted(np.cumsum(closest, dtype=np.float64), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        dist = euclidean_distances(X[cand], X, Y_norm_squared=x2, squared=True)
        np.minimum(closest, dist, out=dist)
        pots = dist.sum(axis=1)
        best = int(np.argmin(pots))
        pot, closest, ids[c] = pots[best], dist[best], cand[best]
    return ids


def run_soft_kmeans(adata_ref, ref_basis="X_pca", K=None, ref_basis_loadings="PCs", sigma=0.1, random_state=None):
    """`_utils.py:311-352` with the reference's estimator call (random_state is ours, for repeatable tests)."""
    from sklearn.cluster import KMeans

    N = adata_ref.shape[0]
    if K is None:
        K = np.min([np.round(N / 30.0), 100]).astype(int)
    model = KMeans(n_clusters=K, init="k-means++", n_init=10, max_iter=25, random_state=random_state)
    model.fit(adata_ref.obsm[ref_basis])
    adata_ref.uns["harmony"] = soft_kmeans_summary(adata_ref.obsm[ref_basis], model.cluster_centers_, sigma,
                                                   ref_basis, ref_basis_loadings)
    return model.inertia_


# --------------------------------------------------------------------------- kNN transfer
def transfer_labels_knn(adata_query, adata_ref, ref_labels, *knn_args, query_labels=None,
                        ref_basis="X_pca_harmony", query_basis="X_pca_harmony", **knn_kwargs):
    """`tl.py:400-436`: the estimator call sequence, verbatim in meaning."""
    from sklearn.neighbors import KNeighborsClassifier

    knn = KNeighborsClassifier(*knn_args, **knn_kwargs)
    knn.fit(adata_ref.obsm[ref_basis], adata_ref.obs[ref_labels])
    if query_labels is None:
        query_labels = ref_labels
    pred = knn.predict(adata_query.obsm[query_basis])
    if isinstance(query_labels, list) and len(query_labels) == 1:
        pred = pred.reshape(pred.shape[0], 1)
    adata_query.obs[query_labels] = pred
    return knn


def knn_indices(ref_emb, query_emb, k):
    """Neighbour indices / squared distances as sklearn's brute Euclidean ArgKmin gives
    them (ascending distance), for index-level parity checks."""
    from sklearn.neighbors import NearestNeighbors

    nn = NearestNeighbors(n_neighbors=k, algorithm="brute").fit(ref_emb)
    dist, idx = nn.kneighbors(query_emb)
    return idx, dist ** 2


def exact_sqdist(ref_emb, query_emb, idx):
    """float64 direct-form squared distances of given (query, neighbour) pairs."""
    diff = np.asarray(query_emb, dtype=np.float64)[:, None, :] - np.asarray(ref_emb, dtype=np.float64)[idx]
    return np.einsum("nkd,nkd->nk", diff, diff)


def knn_vote_from_indices(idx, ref_label_codes, n_classes):
    """sklearn `predict_proba`/`predict` rule for uniform weights:
    proba = votes / k; label = lowest class index among the maxima.  Returns
    (winning class code [N], winning vote fraction [N])."""
    votes = np.zeros((idx.shape[0], n_classes), dtype=np.int64)
    rows = np.repeat(np.arange(idx.shape[0]), idx.shape[1])
    np.add.at(votes, (rows, ref_label_codes[idx].ravel()), 1)
    win = votes.argmax(axis=1)
    return win, votes.max(axis=1) / idx.shape[1]
