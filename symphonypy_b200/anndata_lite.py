"""A minimal AnnData-compatible container.

The drop-in API (`symphonypy_b200.tl`) accepts any object exposing the AnnData
surface the reference's hot path touches (reference
`symphonypy/tl.py:313-397,425-436`, `symphonypy/_utils.py:263-308`):
``X, obs, var, obsm, varm, uns, var_names, obs_names, shape, __len__`` and
``adata[:, <Index of gene names>]``.  Real ``anndata.AnnData`` objects satisfy
this; `AnnDataLite` is provided because `anndata` is not a dependency of this
package (and is not installed in the build image).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class AnnDataLite:
    """Dense-or-CSR expression matrix plus obs/var annotation frames."""

    def __init__(self, X, obs=None, var=None, obsm=None, varm=None, uns=None):
        self.X = X
        n_obs, n_vars = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n_obs).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(n_vars).astype(str))
        assert len(self.obs) == n_obs, "obs must have one row per cell"
        assert len(self.var) == n_vars, "var must have one row per gene"
        self.obsm = dict(obsm) if obsm is not None else {}
        self.varm = dict(varm) if varm is not None else {}
        self.uns = dict(uns) if uns is not None else {}

    @property
    def var_names(self) -> pd.Index:
        return self.var.index

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    @property
    def shape(self):
        return tuple(self.X.shape)

    @property
    def n_obs(self) -> int:
        return self.X.shape[0]

    @property
    def n_vars(self) -> int:
        return self.X.shape[1]

    def __len__(self) -> int:
        return self.X.shape[0]

    def __getitem__(self, key):
        """Supports ``adata[:, genes]`` with a list / Index of gene names (the only indexing form the hot path
        uses, `_utils.py:242,284`) and ``adata[cells, :]`` with positions, a mask or obs names (`_utils.py:431`)."""
        if not (isinstance(key, tuple) and len(key) == 2):
            raise IndexError("AnnDataLite supports adata[rows, genes] indexing only")
        rows, cols = key
        if isinstance(cols, slice):
            col_idx = np.arange(self.n_vars)[cols]
        else:
            cols = pd.Index(cols)
            if cols.dtype == bool:
                col_idx = np.flatnonzero(np.asarray(cols))
            else:
                col_idx = self.var.index.get_indexer(cols)
                if (col_idx < 0).any():
                    raise KeyError("some requested genes are not in var_names")
        if isinstance(rows, slice):
            row_idx = np.arange(self.n_obs)[rows]
        else:
            row_idx = np.asarray(rows)
            if row_idx.dtype == bool:
                row_idx = np.flatnonzero(row_idx)
            elif row_idx.dtype.kind not in "iu":  # obs names (`_utils.py:431` slices a cluster this way)
                row_idx = self.obs.index.get_indexer(pd.Index(rows))
                if (row_idx < 0).any():
                    raise KeyError("some requested cells are not in obs_names")
        X = self.X[row_idx][:, col_idx]
        return AnnDataLite(
            X,
            obs=self.obs.iloc[row_idx],
            var=self.var.iloc[col_idx],
            obsm={k: np.asarray(v)[row_idx] for k, v in self.obsm.items()},
            varm={k: np.asarray(v)[col_idx] for k, v in self.varm.items()},
            uns=self.uns,
        )
