"""
In-tree build of libvegasflow_b200.so (CUDA kernels + C ABI) for sm_100a.

    python -m vegasflow_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  -fmad=false keeps every multiply/add on the
reference arithmetic path separately rounded (explicit fma() calls are kept).
"""
import argparse
import fcntl
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
SO_PATH = os.path.join(LIBDIR, "libvegasflow_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["vf_abi.cu", "vf_aux.cu", "vf_int_symgauss.cu", "vf_int_product.cu", "vf_int_me.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--extended-lambda",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    out = [os.path.join(INCLUDE, "vegasflow_b200.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".cu")):
            out.append(os.path.join(CSRC, f))
    return out


HASH_PATH = os.path.join(LIBDIR, "libvegasflow_b200.sha256")


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive copies)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sorted(_deps()):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(SO_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO_PATH
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU on the same tree)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return SO_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = _nvcc()
    digest = source_hash()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = SO_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, SO_PATH)  # atomic: a concurrent loader never sees a half-written library
    with open(HASH_PATH, "w") as fh:
        fh.write(digest + "\n")
    return SO_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
