"""symphonypy_b200 — B200-native query mapping, a drop-in for the hot path of symphonypy.

    import symphonypy_b200 as sp
    sp.tl.map_embedding(adata_query, adata_ref, key="donor")
    sp.tl.transfer_labels_kNN(adata_query, adata_ref, "cell_type", n_neighbors=10)

Same function names, arguments, `obsm` / `obs` keys and error behaviour as
`symphonypy.tl.map_embedding` (`symphonypy/tl.py:271-397`) and
`symphonypy.tl.transfer_labels_kNN` (`symphonypy/tl.py:400-436`); all arithmetic runs in
hand-written CUDA kernels for sm_100a behind a C ABI (`include/symphony_b200.h`).
Everything else in symphonypy (reference building, confidence metrics, ingest, tsne,
datasets) is out of scope here.
"""
from . import tl  # noqa: F401
from .anndata_lite import AnnDataLite  # noqa: F401

__version__ = "0.1.0"
