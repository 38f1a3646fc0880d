"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a.

    python -m symphonypy_b200.build [--force]

Produces `symphonypy_b200/_C/libsymphony_b200.so` (git-ignored; travels to the GPU box with
the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libsymphony_b200.so")
SOURCES = ["abi.cu", "project.cu", "project_tc.cu", "assign.cu", "moe.cu", "knn.cu", "knn_tc.cu", "confidence.cu", "kmeans.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("SYM_NVCC_EXTRA", "").split()


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, variant: str = "", extra: list[str] | None = None) -> str:
    """Builds the product library; with `variant` (experiments only: instrumented or tuning builds that
    `SYM_B200_LIB` points `_lib.load()` at) a second library `libsymphony_b200_<variant>.so` with `extra` nvcc flags."""
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = LIB if not variant else LIB.replace(".so", f"_{variant}.so")
    flags = FLAGS + list(extra or [])
    stamp = os.path.join(OUT_DIR, f"build{'_' + variant if variant else ''}.stamp")
    dig = _digest() + hashlib.sha256(" ".join(flags).encode()).hexdigest()
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    obj_dir = OUT_DIR if not variant else os.path.join(OUT_DIR, variant)
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [NVCC, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(obj_dir, src + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-4000:]}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([NVCC, "-shared", "-o", lib, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


if __name__ == "__main__":
    # python -m symphonypy_b200.build [--force] [-v] [--variant NAME -DFLAG ...]
    argv = sys.argv[1:]
    var = argv[argv.index("--variant") + 1] if "--variant" in argv else ""
    print(build(force="--force" in argv, verbose="-v" in argv, variant=var, extra=[a for a in argv if a.startswith("-D")]))
