"""CPU oracle for the symphonypy query-mapping path.  TEST INFRASTRUCTURE ONLY.

Nothing under `symphonypy_b200/` imports this package.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it, and only as the checker or the timed CPU baseline.

Parity status: PINNED.  The numpy restatement in `symphony_oracle.py` is checked
(tests/test_oracle.py) against golden vectors in `tests/golden/` that were
produced by running the UNMODIFIED reference functions (`/root/reference/
symphonypy/tl.py`, `_utils.py`, imported through `ref_loader.py`) on seeded
synthetic inputs with `gen_golden.py`; when `/root/reference` is present the same
tests also run the reference live, side by side.  The reference ships no golden
vectors of its own for this path (its `tests/data/*.h5ad` are absent from the
checkout).
"""
