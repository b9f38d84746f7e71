"""In-tree build of libeoe_b200.so (hand-written sm_100a CUDA + the C ABI of include/eoe_b200.h).

`python -m eoe_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU; the .so is
git-ignored but travels with the repo snapshot to the GPU box.  There is deliberately no CPU fallback:
`eoe_b200._lib.lib()` raises if the library is missing.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libeoe_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "eoe_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def source_id() -> str:
    """sha256 over csrc/* and include/eoe_b200.h: the identity of the kernels a library was built from.  Baked into the
    library (eoe_build_id()) so that profile artefacts (profiles/gemm_traffic.json) can be tied to the build they measured."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(os.path.dirname(HERE), "include", "eoe_b200.h")]:
        h.update(os.path.basename(path).encode() + b"\0")
        h.update(open(path, "rb").read())
    return h.hexdigest()[:16]


GEMM_SOURCES = ("gemm_sm100.cuh", "sm100_ptx.cuh", "common.cuh", "attention_sm100.cuh", "vit.cu")


def gemm_source_id() -> str:
    """sha256 over the sources the encoder's GEMM / attention kernels and their launches are compiled from (and the nvcc
    flags): an ncu capture of those kernels (profiles/gemm_traffic.json) stays valid across builds that only touch the
    heads / AUC translation units, and only across those."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in GEMM_SOURCES + ("../../include/eoe_b200.h",):
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if os.path.basename(src) == "capi.cu":       # carries the build id: always rebuilt with the link (a 1 s compile)
        cmd = [_nvcc(), *NVCC_FLAGS, f'-DEOE_BUILD_ID="{source_id()}"', "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr
    hdr_m = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hdr_m = max(hdr_m, os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "eoe_b200.h")))
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_m):
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(obj[:-2] + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile, sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
