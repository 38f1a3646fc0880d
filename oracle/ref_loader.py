"""Import the unmodified reference (`/root/reference/symphonypy`) for validation.

The reference imports `anndata`, `scanpy` and `harmonypy` at module top
(`symphonypy/tl.py:12,16`, `_utils.py:10,11,15`) although the query-mapping path
never uses them, and none of the three is installed here.  Stub modules are
planted in `sys.modules`, the package `__init__` (which pulls in `datasets` ->
`scanpy.read`) is bypassed, and `symphonypy.tl` / `symphonypy._utils` are then
imported as they are.  Nothing is copied; the loader only points at the tree.

`/root/reference` exists in the build container only — callers must check
`available()` and never rely on it on the GPU box.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SYMPHONY_REFERENCE_ROOT", "/root/reference")
_ALIAS = "symphonypy"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "symphonypy", "tl.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__dict__["__symphony_stub__"] = True
        sys.modules[name] = m
    for k, v in attrs.items():
        if not hasattr(m, k):
            setattr(m, k, v)
    return m


def load():
    """Returns (tl, _utils) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if _ALIAS in sys.modules and getattr(sys.modules[_ALIAS], "__symphony_ref__", False):
        return sys.modules[_ALIAS + ".tl"], sys.modules[_ALIAS + "._utils"]

    from symphonypy_b200.anndata_lite import AnnDataLite

    try:
        import anndata  # noqa: F401
    except ImportError:
        _stub("anndata", AnnData=AnnDataLite)
    try:
        import harmonypy  # noqa: F401
    except ImportError:
        def run_harmony(*a, **k):
            raise NotImplementedError("harmonypy is not installed (reference building is out of scope)")
        _stub("harmonypy", run_harmony=run_harmony)
    try:
        import scanpy  # noqa: F401
    except ImportError:
        class Ingest:  # placeholder base class for `Ingest_sp` (`_utils.py:355`)
            def __init__(self, *a, **k):
                pass
        sc = _stub("scanpy")
        tools = _stub("scanpy.tools")
        ing = _stub("scanpy.tools._ingest", Ingest=Ingest, _DimDict=dict)
        compat = _stub("scanpy._compat", pkg_version=lambda name: None)
        sc.tools, sc._compat, tools._ingest = tools, compat, ing

    pkg = types.ModuleType(_ALIAS)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "symphonypy")]
    pkg.__symphony_ref__ = True
    sys.modules[_ALIAS] = pkg
    utils = importlib.import_module(_ALIAS + "._utils")
    tl = importlib.import_module(_ALIAS + ".tl")
    return tl, utils
