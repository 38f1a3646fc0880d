"""CPU tests of the host-side mirror of the reference interface (no kernels involved)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from oracle import symphony_oracle as so
from symphonypy_b200 import _design, _lib, synth, tl
from symphonypy_b200.anndata_lite import AnnDataLite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _query_with_obs(n=200, seed=0):
    rng = np.random.default_rng(seed)
    obs = pd.DataFrame({
        "donor": rng.choice(["b_10", "b_2", "b_9", "a"], n),   # string order != numeric order
        "chem": rng.choice([3, 10, 1], n),                      # ints are str()-ed by the reference
        "site": pd.Categorical(rng.choice(["x", "y"], n)),
    })
    return AnnDataLite(np.zeros((n, 3), dtype=np.float32), obs=obs)


@pytest.mark.parametrize("key", [None, "donor", ["donor"], ["chem", "donor", "site"]])
def test_design_codes_match_get_dummies(key):
    q = _query_with_obs()
    codes, levels = _design.batch_codes(q, key)
    phi_, lam = so.build_design(q, key, None)
    n_levels = [len(lv) for lv in levels]
    assert phi_.shape[0] == 1 + sum(n_levels)
    # rebuild the one-hot matrix from the codes: must equal the reference's Phi_ row for row
    mine = np.zeros_like(phi_)
    mine[0] = 1
    off = 1
    for v, lv in enumerate(n_levels):
        mine[off + codes[v], np.arange(codes.shape[1])] = 1
        off += lv
    np.testing.assert_array_equal(mine, phi_)
    np.testing.assert_array_equal(np.diag(lam), _design.ridge_diagonal(None, n_levels))


@pytest.mark.parametrize("lamb", [None, 0.5, 2, [0.5, 2.0, 3.0], list(np.arange(1, 10) / 4)])
def test_ridge_diagonal_matches_reference(lamb):
    q = _query_with_obs()
    key = ["chem", "donor", "site"]  # 3 + 4 + 2 = 9 levels
    _, levels = _design.batch_codes(q, key)
    _, lam = so.build_design(q, key, lamb if lamb is None or np.ndim(lamb) == 0 else list(lamb))
    np.testing.assert_allclose(np.diag(lam), _design.ridge_diagonal(lamb, [len(l) for l in levels]))


def test_ridge_diagonal_bad_length():
    with pytest.raises(AssertionError, match="each batch variable must have a lambda"):
        _design.ridge_diagonal([1.0, 2.0], [3, 4, 2])


def test_knn_params_surface():
    assert tl._knn_params((), {})["n_neighbors"] == 5
    assert tl._knn_params((7,), {"algorithm": "kd_tree", "n_jobs": 4})["n_neighbors"] == 7
    assert tl._knn_params((), {"metric": "euclidean", "n_neighbors": 3})["n_neighbors"] == 3
    assert tl._knn_params((), {"weights": "distance"})["weights"] == "distance"
    with pytest.raises(NotImplementedError):
        tl._knn_params((), {"weights": lambda d: d})
    with pytest.raises(NotImplementedError):
        tl._knn_params((), {"metric": "cosine"})
    with pytest.raises(NotImplementedError):
        tl._knn_params((), {"p": 1})
    with pytest.raises(ValueError):
        tl._knn_params((0,), {})
    with pytest.raises(TypeError):
        tl._knn_params((), {"neighbours": 3})


def test_map_embedding_asserts_like_reference():
    q, r = synth.small_case(N=20, M=50, G=32, d=20, K=4, L=3, seed=1)
    bad = AnnDataLite(r.X, obs=r.obs, var=r.var.drop(columns=["mean"]), obsm=r.obsm, varm=r.varm, uns=r.uns)
    with pytest.raises(AssertionError, match="Gene expression means"):
        tl.map_embedding(q, bad)
    with pytest.raises(AssertionError, match="not found in adata_ref.var"):
        tl.map_embedding(q, r, use_genes_column="nope")


def test_missing_harmony_warns_like_reference_then_needs_the_gpu(monkeypatch):
    """`tl.py:320-335`: the k-means fallback is announced with the reference's warning; it runs on the GPU only."""
    import torch

    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    q, r = synth.small_case(N=20, M=50, G=32, d=20, K=4, L=3, seed=1)
    no_h = AnnDataLite(r.X, obs=r.obs, var=r.var, obsm=r.obsm, varm=r.varm, uns={})
    with pytest.warns(UserWarning, match="Not found `harmony` object in adata_ref.uns"):
        with pytest.raises(_lib.SymphonyLibraryError, match="no CPU path"):
            tl.map_embedding(q, no_h)


def test_no_cpu_fallback(monkeypatch):
    """Without a CUDA device the product path refuses to run; it never routes through the oracle."""
    import torch

    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    q, r = synth.small_case(N=20, M=50, G=32, d=20, K=4, L=3, seed=1)
    with pytest.raises(_lib.SymphonyLibraryError, match="no CPU path"):
        tl.map_embedding(q, r)
    with pytest.raises(_lib.SymphonyLibraryError, match="no CPU path"):
        tl.transfer_labels_kNN(q, r, "cell_type")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "symphonypy_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), name


def test_library_exports_every_declared_symbol(built_library):
    header = open(os.path.join(ROOT, "include", "symphony_b200.h")).read()
    declared = set(re.findall(r"\b(sym_[a-z0-9_]+)\s*\(", header))
    declared -= {"sym_stream_t"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(built_library)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.sym_version() == 100 and loaded.sym_compiled_arch() == 100
    assert loaded.sym_last_error() == b""


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SymphonyLibraryError, match="no CPU fallback"):
        _lib.load()


def test_anndata_lite_gene_indexing():
    X = np.arange(12, dtype=np.float32).reshape(3, 4)
    a = AnnDataLite(X, var=pd.DataFrame(index=pd.Index(["a", "b", "c", "d"])))
    sub = a[:, pd.Index(["d", "b"])]
    np.testing.assert_array_equal(sub.X, X[:, [3, 1]])
    assert list(sub.var_names) == ["d", "b"] and len(a) == 3


def test_decode_labels_matches_plain_numpy_take():
    """The Arrow dictionary decode used for string labels builds the same column (values and dtype) as assigning
    sklearn's object array of predictions, and non-string classes take the numpy path."""
    import pandas as pd

    rng = np.random.default_rng(0)
    classes = np.array(["B cell", "T cell", "ü-type", ""], dtype=object)
    lab = rng.integers(0, len(classes), 1000).astype(np.int32)
    idx = pd.RangeIndex(1000).astype(str)
    obs = pd.DataFrame(index=idx)
    obs["fast"] = tl._decode_labels(classes, lab, idx)
    obs["plain"] = classes[lab]
    assert obs["fast"].dtype == obs["plain"].dtype
    assert obs["fast"].equals(obs["plain"])
    ints = tl._decode_labels(np.array([3, 7, 11]), lab % 3, idx)
    assert isinstance(ints, np.ndarray) and ints.dtype.kind == "i" and (ints == np.array([3, 7, 11])[lab % 3]).all()


def test_knn_tier_logic_with_a_fake_scan(monkeypatch):
    """`engine.knn_search_certified` on CPU tensors with the scan replaced by a stub: uncertified rows go
    FP16 -> BF16x3 -> FP32, only the flagged rows are re-run, and an index that fails the FP16 certificate too
    often starts with BF16x3 from then on."""
    import torch

    from symphonypy_b200 import engine

    class FakeLib:
        def sym_knn_method_supported(self, n, M, d, k, m):
            return 1

    monkeypatch.setattr(engine._lib, "load", lambda: FakeLib())
    calls = []

    def fake_search(index, Q, k, method=engine.KNN_AUTO, want_unsafe=True, events=None, prune=False):
        n = Q.shape[0]
        calls.append((method, n))
        tag = Q[:, 0].long()                      # row identity travels in column 0
        idx = (tag[:, None] * 10 + method).int().repeat(1, k)
        dist2 = torch.zeros((n, k))
        # FP16 fails the rows whose tag is a multiple of `index.fail16`, BF16x3 those that are multiples of 50
        if method == engine.KNN_TC16:
            unsafe = (tag % index.fail16 == 0).to(torch.uint8)
        elif method == engine.KNN_TC:
            unsafe = (tag % 50 == 0).to(torch.uint8)
        else:
            unsafe = torch.zeros(n, dtype=torch.uint8)
        return idx, dist2, unsafe

    monkeypatch.setattr(engine, "knn_search", fake_search)
    N, k = 5000, 3
    Q = torch.arange(N, dtype=torch.float32)[:, None].repeat(1, 4)
    index = engine.KnnIndex(ref=torch.zeros((10, 4)), packed=torch.zeros(1, dtype=torch.uint8), M=10, d=4)

    index.fail16 = 25                             # 4 % of the rows fail the FP16 certificate: FP16 stays first
    idx, dist2, n_rep = engine.knn_search_certified(index, Q, k)
    assert calls == [(engine.KNN_TC16, N), (engine.KNN_TC, 200), (engine.KNN_FFMA, 100)]
    assert n_rep == 200 and k not in index.first_tier
    tier = idx[:, 0] % 10
    rows = torch.arange(N)
    assert (tier[rows % 25 != 0] == engine.KNN_TC16).all()
    assert (tier[(rows % 25 == 0) & (rows % 50 != 0)] == engine.KNN_TC).all()
    assert (tier[rows % 50 == 0] == engine.KNN_FFMA).all()
    assert (idx[:, 0] // 10 == rows).all()        # every row got its own result back

    calls.clear()
    index.fail16 = 5                              # 20 % fail: the index remembers and starts with BF16x3 next time
    engine.knn_search_certified(index, Q, k)
    assert calls[0] == (engine.KNN_TC16, N) and index.first_tier[k] == engine.KNN_TC
    calls.clear()
    engine.knn_search_certified(index, Q, k)
    assert calls == [(engine.KNN_TC, N), (engine.KNN_FFMA, 100)]

    calls.clear()                                 # an explicit method is never adapted
    engine.knn_search_certified(index, Q, k, engine.KNN_TC16)
    assert calls[0] == (engine.KNN_TC16, N)

    # a handful of uncertified rows: exact brute force first (bounded by the first tier's k-th distance), and only the
    # rows it could not finish go on to the scan tiers
    exact_calls = []

    def fake_exact(index, Q, rows, k, ub):
        exact_calls.append((rows.tolist(), ub.tolist()))
        tag = Q[rows, 0].long()
        over = (tag % 200 == 0).to(torch.uint8)   # "bound useless" for these
        return (tag[:, None] * 10 + 9).int().repeat(1, k), torch.ones((len(rows), k)), over

    monkeypatch.setattr(engine, "knn_exact_rows", fake_exact)
    calls.clear()
    index2 = engine.KnnIndex(ref=torch.zeros((10, 4)), packed=torch.zeros(1, dtype=torch.uint8), M=10, d=4)
    index2.fail16 = 100                           # 50 rows fail
    idx, dist2, n_rep = engine.knn_search_certified(index2, Q, k)
    assert n_rep == 50 and len(exact_calls) == 1 and exact_calls[0][0] == list(range(0, N, 100))
    assert all(u > 0 for u in exact_calls[0][1])  # the k-th distance (0 here) nudged up
    assert calls == [(engine.KNN_TC16, N), (engine.KNN_TC, 25), (engine.KNN_FFMA, 25)]
    tier = idx[:, 0] % 10
    assert (tier[(rows % 100 == 0) & (rows % 200 != 0)] == 9).all() and (tier[rows % 200 == 0] == engine.KNN_FFMA).all()
    assert (idx[:, 0] // 10 == rows).all()


def test_cluster_maha_dists_matches_oracle():
    """Host half of `per_cluster_confidence`: from exact per-cluster moments it reproduces the oracle's
    (= the reference's) distances, NaN for clusters below u*d cells included."""
    from oracle import symphony_oracle as so

    rng = np.random.default_rng(3)
    d, K = 6, 9
    cent = rng.normal(size=(K, d)) * 5
    cent_norm = cent / np.linalg.norm(cent, axis=1, keepdims=True)
    groups = [rng.normal(size=(n, d)) @ rng.normal(size=(d, d)) + rng.normal(size=d) * 3 for n in (40, 13, 5, 200)]
    size = np.array([g.shape[0] for g in groups])
    mean = np.stack([g.mean(0) for g in groups])
    scatter = np.stack([(g - g.mean(0)).T @ (g - g.mean(0)) for g in groups])
    for u, lamb in ((2, 0.0), (1, 0.3)):
        got = tl._cluster_maha_dists(size, mean, scatter, cent, cent_norm, u, lamb)
        want = [so.cluster_maha_dist(g, cent, cent_norm, u, lamb) for g in groups]
        want = np.array([np.nan if w is None else w for w in want])
        np.testing.assert_allclose(got, want, rtol=1e-10, equal_nan=True)
        assert np.isnan(got).sum() == (size < u * d).sum()


def test_public_header_is_plain_c(tmp_path):
    """`include/symphony_b200.h` is the drop-in boundary: it has to compile as C99 and as C++ with nothing but
    the standard headers (no torch, no CUDA types), and a C program has to link against the built library."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_header.c"
    src.write_text('#include "symphony_b200.h"\n'
                   'int main(void) { return sym_version() > 0 && sym_compiled_arch() == 100 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                   check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)],
                   check=True)


def test_c_program_links_and_calls_the_library(tmp_path, built_library):
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "link.c"
    src.write_text('#include <stdio.h>\n#include "symphony_b200.h"\n'
                   'int main(void) { printf("%d %d\\n", sym_version(), sym_compiled_arch()); '
                   'return sym_knn_method_supported(1000, 5000, 30, 10, 3) == 1 ? 0 : 2; }\n')
    exe = tmp_path / "link"
    libdir = os.path.dirname(built_library)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lsymphony_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) > 0 and int(out[1]) == 100   # host-only entry points: no GPU needed


# ------------------------------------------------------------------------------ round 2: host logic without N-row string work
def test_design_codes_categorical_with_unused_and_missing_levels():
    """Categorical columns go through their codes: unused categories give no column, int categories sort as
    strings; a missing value becomes its own level "nan" (pandas < 3 `astype(str)` semantics for NaN)."""
    rng = np.random.default_rng(3)
    n = 300
    obs = pd.DataFrame({
        "a": pd.Categorical(rng.choice(["u", "v"], n), categories=["z_unused", "v", "u"]),
        "b": pd.Categorical(rng.choice([10, 2, 9], n)),
        "c": pd.Series(rng.choice(["p", "q", None], n), dtype=object),
        "d": pd.Categorical(rng.choice(["m", "n", None], n)),
    })
    q = AnnDataLite(np.zeros((n, 2), dtype=np.float32), obs=obs)
    codes, levels = _design.batch_codes(q, ["c", "d"])
    assert levels == [["nan", "p", "q"], ["m", "n", "nan"]]
    assert (codes[0] == 0).sum() == obs["c"].isna().sum() and (codes[1] == 2).sum() == obs["d"].isna().sum()
    for key in ("a", "b", ["b", "a"]):
        codes, levels = _design.batch_codes(q, key)
        phi_, _ = so.build_design(q, key, None)
        n_levels = [len(lv) for lv in levels]
        mine = np.zeros_like(phi_)
        mine[0] = 1
        off = 1
        for v, lv in enumerate(n_levels):
            mine[off + codes[v], np.arange(n)] = 1
            off += lv
        np.testing.assert_array_equal(mine, phi_)


def test_design_codes_never_stringify_all_rows(monkeypatch):
    """1M cells: the codes come from factorize / cat.codes; `Series.astype(str)` only ever sees the distinct values."""
    n = 1_000_000
    rng = np.random.default_rng(0)
    obs = pd.DataFrame({"batch": pd.Categorical.from_codes(rng.integers(0, 100, n), [f"b{i}" for i in range(100)]),
                        "num": rng.integers(0, 7, n)})
    q = AnnDataLite(np.zeros((n, 1), dtype=np.float32), obs=obs)
    longest = [0]
    orig = pd.Series.astype

    def spy(self, dtype, *a, **k):
        if dtype is str:
            longest[0] = max(longest[0], len(self))
        return orig(self, dtype, *a, **k)

    monkeypatch.setattr(pd.Series, "astype", spy)
    codes, levels = _design.batch_codes(q, ["batch", "num"])
    assert longest[0] <= 100
    assert levels[0] == sorted(f"b{i}" for i in range(100)) and levels[1] == [str(i) for i in range(7)]
    np.testing.assert_array_equal(np.asarray(levels[0], dtype=object)[codes[0]], np.asarray(obs["batch"].astype(object)))
    np.testing.assert_array_equal(codes[1], obs["num"].to_numpy())


def test_content_digest_notices_in_place_edits():
    a = np.random.default_rng(0).standard_normal((300_000, 30))          # 72 MB: goes through the threaded path
    d0 = tl.content_digest(a)
    assert tl.content_digest(a) == d0 and tl.content_digest(a.copy()) == d0
    a[123_456, 7] += 1e-9
    d1 = tl.content_digest(a)
    assert d1 != d0
    a[[0, 1]] = a[[1, 0]]                                                   # a row swap is a change too
    assert tl.content_digest(a) != d1
    assert tl.content_digest(a.astype(np.float32)) != tl.content_digest(a)
    assert tl.content_digest(a[:, :15]) == tl.content_digest(np.ascontiguousarray(a[:, :15]))
    assert tl.content_digest(None, np.arange(3)) != tl.content_digest(np.arange(3), None)
    lab = np.array(["x", "y", "x"], dtype=object)
    e0 = tl.content_digest(lab)
    lab[1] = "z"
    assert tl.content_digest(lab) != e0


def test_host_digest_is_thread_count_independent_and_sees_every_byte(built_library):
    """`sym_host_digest64` (host code of the C-ABI library, no GPU): the same value on 1, 3 and 8 threads, a different
    one after flipping any single bit (block interior, the zero-padded tail, a chunk boundary) or appending a zero."""
    from symphonypy_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(1)
    buf = rng.integers(0, 256, (20 << 20) + 37, dtype=np.uint8)             # three 8 MiB chunks, ragged tail

    def dig(b, threads=4):
        return int(lib.sym_host_digest64(b.ctypes.data if b.size else None, b.nbytes, threads))

    base = dig(buf)
    assert {dig(buf, t) for t in (1, 3, 8, 0)} == {base}
    for pos in (0, 63, 64, (8 << 20) - 1, 8 << 20, (16 << 20) + 5, buf.size - 1):
        buf[pos] ^= 0x10
        assert dig(buf) != base, pos
        buf[pos] ^= 0x10
    assert dig(buf) == base
    small = np.zeros(70, dtype=np.uint8)
    seen = {dig(small[:n]) for n in range(0, 71)}                           # all-zero inputs of every length differ
    assert len(seen) == 71
    assert dig(np.zeros((8 << 20) + 1, np.uint8)) != dig(np.zeros(8 << 20, np.uint8))


def test_host_copy_and_staging_with_one_intra_op_thread(built_library, monkeypatch):
    """`sym_host_copy` (threaded memcpy of the C-ABI library, host only) copies exactly the bytes it is given, and
    `tl._stage_copy` takes it for same-dtype copies when the process has fewer torch intra-op threads than its share of
    the cores (torchrun's OMP_NUM_THREADS=1) — dtype conversions still go through torch."""
    import torch

    from symphonypy_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(2)
    for n in (0, 1, 4093, (4 << 20) - 1, 4 << 20, (9 << 20) + 5):
        src = rng.integers(0, 256, n, dtype=np.uint8)
        for threads in (1, 3, 8):
            dst = np.full(n + 16, 7, dtype=np.uint8)
            lib.sym_host_copy(dst.ctypes.data + 8, src.ctypes.data, n, threads)
            assert (dst[8:8 + n] == src).all() and (dst[:8] == 7).all() and (dst[8 + n:] == 7).all(), (n, threads)

    calls = []
    real = lib.sym_host_copy

    class Spy:
        def __getattr__(self, name):
            return getattr(lib, name)

        def sym_host_copy(self, *a):
            calls.append(a[2])
            return real(*a)

    monkeypatch.setattr(tl._lib, "load", lambda: Spy())
    monkeypatch.setattr(tl, "_hash_threads", lambda: 4)
    monkeypatch.setattr(torch, "get_num_threads", lambda: 1)
    src = rng.standard_normal(600_000).astype(np.float32)
    dst = torch.zeros(600_000, dtype=torch.float32)
    tl._stage_copy(dst, src)
    assert calls == [src.nbytes] and np.array_equal(dst.numpy(), src)
    idx = rng.integers(0, 2000, 500_000).astype(np.int32)
    dsti = torch.zeros(500_000, dtype=torch.int32)
    tl._stage_copy(dsti, idx)
    assert len(calls) == 2 and np.array_equal(dsti.numpy(), idx)
    src64 = rng.standard_normal(400_000)                                  # float64 -> float32: torch converts
    tl._stage_copy(dst[:400_000], src64)
    assert len(calls) == 2 and np.array_equal(dst[:400_000].numpy(), src64.astype(np.float32))
    tl._stage_copy(dst[:1000], src[:1000])                                # small: not worth threads
    assert len(calls) == 2
    monkeypatch.setattr(torch, "get_num_threads", lambda: 16)             # torch has the cores: its own copy
    tl._stage_copy(dst, src)
    assert len(calls) == 2 and np.array_equal(dst.numpy(), src)


def test_decode_labels_in_chunks(monkeypatch):
    """Long label columns are gathered on a few threads; the column equals the single-call one."""
    import pandas as pd

    rng = np.random.default_rng(3)
    classes = np.array([f"type {i}" for i in range(37)], dtype=object)
    lab = rng.integers(0, 37, 5003).astype(np.int32)
    idx = pd.RangeIndex(5003).astype(str)
    whole = tl._decode_labels(classes, lab, idx)
    monkeypatch.setattr(tl, "_DECODE_CHUNK", 512)
    parts = tl._decode_labels(classes, lab, idx)
    obs = pd.DataFrame(index=idx)
    obs["parts"], obs["whole"], obs["plain"] = parts, whole, classes[lab]
    assert obs["parts"].dtype == obs["plain"].dtype and obs["parts"].equals(obs["plain"]) and obs["whole"].equals(obs["plain"])


@pytest.mark.parametrize("kind", ["category", "str", "object", "int", "category_unused", "bool"])
def test_encode_labels_equals_numpy_unique(kind):
    rng = np.random.default_rng(1)
    n = 5000
    if kind in ("int",):
        col = pd.Series(rng.choice([7, -3, 100, 12], n))
    elif kind == "bool":
        col = pd.Series(rng.choice([True, False], n))
    else:
        vals = rng.choice(["T cell", "B cell", "b cell", "NK", "10", "2"], n)
        col = pd.Series(vals, dtype=object if kind == "object" else None)
        if kind == "category":
            col = col.astype("category")
        if kind == "category_unused":
            col = pd.Series(pd.Categorical(vals, categories=["zz", "NK", "10", "2", "b cell", "B cell", "T cell", "aa"]))
    classes, codes = tl._encode_labels(col, "cpu")
    want_classes, want_inv = np.unique(np.asarray(col), return_inverse=True)   # sklearn: classes_ = np.unique(y)
    assert list(classes) == list(want_classes)
    np.testing.assert_array_equal(codes.numpy(), want_inv)


def test_pad_runs_pads_every_run_by_repeating_its_last_row():
    """Host logic of the kNN tile pruning: the sorted query rows are padded so that no group of 32 scan rows straddles
    two clusters (`engine.pad_runs`)."""
    import torch

    from symphonypy_b200 import engine

    perm = torch.tensor([7, 3, 9, 1, 4, 8, 0, 2, 5, 6], dtype=torch.int32)
    seg = torch.tensor([0, 3, 3, 4, 10], dtype=torch.int64)          # runs of 3, 0, 1 and 6 rows
    out = engine.pad_runs(perm, seg, 4).tolist()
    assert out == [7, 3, 9, 9, 1, 1, 1, 1, 4, 8, 0, 2, 5, 6, 6, 6]
    assert engine.pad_runs(perm, seg, 1).tolist() == perm.tolist()
    # every original row appears, padding only repeats rows of the same run
    rng = np.random.default_rng(0)
    code = torch.from_numpy(rng.integers(0, 13, 1000))
    order = torch.argsort(code, stable=True).to(torch.int32)
    counts = torch.bincount(code, minlength=13)
    seg = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)])
    out = engine.pad_runs(order, seg, 32)
    assert out.numel() % 32 == 0 and set(out.tolist()) == set(range(1000))
    groups = code[out.long()].reshape(-1, 32)
    assert bool((groups == groups[:, :1]).all())                    # no group of 32 mixes two runs


def test_csr_row_chunks_cover_all_rows_within_the_byte_budget():
    """Host logic of the CSR streaming (`tl.csr_row_chunks`): consecutive row ranges that cover every row once, hold at
    most the budgeted number of non-zeros unless a single row exceeds it, and never an empty range."""
    from symphonypy_b200 import tl

    rng = np.random.default_rng(3)
    for nnz_per_row, budget in (((rng.poisson(20, 500)), 300), (np.array([0, 0, 5, 1000, 0, 3, 0]), 10), (np.zeros(7, int), 4),
                                (np.array([9]), 100), (np.array([], dtype=int), 5)):
        indptr = np.concatenate([[0], np.cumsum(nnz_per_row)]).astype(np.int64)
        b = tl.csr_row_chunks(indptr, budget)
        N = len(nnz_per_row)
        assert b[0] == 0 and b[-1] == N and all(x < y for x, y in zip(b[:-1], b[1:])) or N == 0 and b == [0]
        for s0, e0 in zip(b[:-1], b[1:]):
            n = indptr[e0] - indptr[s0]
            assert n <= budget or e0 - s0 == 1
