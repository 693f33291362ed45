"""Builds the CUDA library in-tree: rbepwt_b200/_lib/librbepwt_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.environ.get("RBEPWT_B200_LIB") or os.path.join(LIB_DIR, "librbepwt_b200.so")  # env: experiments only

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # fp64 products and sums are never contracted; the one FMA on the path is explicit
    "-Xcompiler", "-fPIC", "-shared",
    "--cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + [
        os.path.join(HERE, "..", "include", "rbepwt_b200.h")]


def needs_build():
    if os.environ.get("RBEPWT_B200_LIB"):
        return False
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force=False, verbose=False):
    """Several processes may get here at once (one rank per GPU under torchrun): one builds, under a file lock, into a
    temporary file that is renamed over the library -- nobody ever dlopens a half-written file -- and the others find
    it up to date when they get the lock."""
    import fcntl

    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
                "-o", tmp, os.path.join(CSRC, "rbepwt_b200.cu")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building librbepwt_b200.so")
            os.replace(tmp, LIB_PATH)
            if verbose:
                sys.stderr.write(res.stdout + res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose="-v" in sys.argv))
