"""Build the C-ABI shared library ``diffsptk_b200/lib/libdiffsptk_b200.so`` in-tree with nvcc.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); no torch headers are involved, so
the library is loadable from any host language through ``include/diffsptk_b200.h``.
"""

from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(ROOT, "build", "obj_" + os.environ.get("DSB200_LIB_NAME", "default"))
LIB_PATH = os.path.join(OUT_DIR, os.environ.get("DSB200_LIB_NAME", "libdiffsptk_b200.so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
    *os.environ.get("DSB200_EXTRA_NVCC_FLAGS", "").split(),
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA toolkit is required to build diffsptk_b200")
    return exe


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "diffsptk_b200.h"))
    stamp = os.path.join(OUT_DIR, ".build_digest_" + os.path.basename(LIB_PATH))
    digest = _digest(deps)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    headers = [d for d in deps if not d.endswith(".cu")]

    def compile_one(src):
        # an object is reused when its source, every header and the flags are unchanged (build/ is scratch: it is
        # neither tracked nor shipped, so a fresh checkout compiles everything)
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        obj_digest = _digest([src] + headers)
        if not force and not verbose and os.path.exists(obj) and os.path.exists(obj + ".digest") \
                and open(obj + ".digest").read() == obj_digest:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(obj + ".digest", "w") as f:
            f.write(obj_digest)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    # link beside the target and rename: a reader (a loader, a snapshot of the tree) never sees a partial library
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
