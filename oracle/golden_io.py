"""Save / load a (query, reference) AnnDataLite pair plus expected outputs as one .npz.

TEST INFRASTRUCTURE.  Fixtures are self-contained (inputs and outputs), so the GPU
box — which has no `/root/reference` — can replay them.
"""
from __future__ import annotations

import json

import numpy as np
import pandas as pd
import scipy.sparse as sp

from symphonypy_b200.anndata_lite import AnnDataLite


def _s(a):
    return np.asarray(a, dtype=str)


def pack(query, ref, call: dict, outputs: dict) -> dict:
    Xq = query.X
    d = {
        "q_sparse": np.array(sp.issparse(Xq)),
        "q_X": Xq.toarray() if sp.issparse(Xq) else np.asarray(Xq),
        "q_var_names": _s(query.var_names),
        "q_obs_names": _s(query.obs_names),
        "q_obs_cols": _s(list(query.obs.columns)),
        "q_uns_log1p": np.array("log1p" in query.uns),
        "r_var_names": _s(ref.var_names),
        "r_obs_names": _s(ref.obs_names),
        "r_mean": np.asarray(ref.var["mean"]),
        "r_std": np.asarray(ref.var["std"]),
        "r_hv": np.asarray(ref.var["highly_variable"]),
        "r_PCs": np.asarray(ref.varm["PCs"]),
        "r_emb": np.asarray(ref.obsm["X_pca_harmony"]),
        "r_C": np.asarray(ref.uns["harmony"]["C"]),
        "r_Nr": np.asarray(ref.uns["harmony"]["Nr"]),
        "r_K": np.array(ref.uns["harmony"]["K"]),
        "r_obs_cols": _s(list(ref.obs.columns)),
        "call": np.array(json.dumps(call)),
    }
    for c in query.obs.columns:
        d[f"q_obs__{c}"] = _s(query.obs[c])
    for c in ref.obs.columns:
        d[f"r_obs__{c}"] = _s(ref.obs[c])
    for k, v in outputs.items():
        d[f"out__{k}"] = np.asarray(v) if np.asarray(v).dtype != object else _s(v)
    return d


def unpack(path):
    z = np.load(path, allow_pickle=False)
    Xq = z["q_X"]
    q_obs = pd.DataFrame({c: z[f"q_obs__{c}"].astype(object) for c in z["q_obs_cols"]},
                         index=pd.Index(z["q_obs_names"].astype(object)))
    query = AnnDataLite(sp.csr_matrix(Xq) if bool(z["q_sparse"]) else Xq, obs=q_obs,
                        var=pd.DataFrame(index=pd.Index(z["q_var_names"].astype(object))),
                        uns={"log1p": {}} if bool(z["q_uns_log1p"]) else {})
    r_obs = pd.DataFrame({c: z[f"r_obs__{c}"].astype(object) for c in z["r_obs_cols"]},
                         index=pd.Index(z["r_obs_names"].astype(object)))
    var = pd.DataFrame({"mean": z["r_mean"], "std": z["r_std"], "highly_variable": z["r_hv"]},
                       index=pd.Index(z["r_var_names"].astype(object)))
    ref = AnnDataLite(sp.csr_matrix((len(r_obs), len(var)), dtype=np.float32), obs=r_obs, var=var,
                      obsm={"X_pca_harmony": z["r_emb"]}, varm={"PCs": z["r_PCs"]},
                      uns={"harmony": {"C": z["r_C"], "Nr": z["r_Nr"], "K": int(z["r_K"]),
                                       "ref_basis_loadings": "PCs"}})
    call = json.loads(str(z["call"]))
    outputs = {k[5:]: z[k] for k in z.files if k.startswith("out__")}
    return query, ref, call, outputs
