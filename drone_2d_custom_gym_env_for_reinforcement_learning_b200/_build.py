"""Builds csrc/libdrone2d.so (sm_100a only) with nvcc.  The .so is kept in-tree
(git-ignored) so it travels with the repository snapshot to the GPU box."""
import fcntl
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdrone2d.so")
SOURCES = ["drone2d.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _inputs():
    """Every file the library is compiled from: the sources, all headers next to them and the public header."""
    files = [os.path.join(CSRC, s) for s in SOURCES]
    files += sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")))
    files += sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    return files


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _inputs())


def build(force=False, verbose=False, extra_flags=()):
    """Compile the CUDA extension if it is missing or older than its sources.  Safe under torchrun: the
    ranks serialise on a file lock, the first one builds into a private temporary file and publishes it
    with an atomic rename, the others find an up-to-date library when they get the lock."""
    if not force and not is_stale():
        return LIB
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            nvcc = _nvcc()
            if nvcc is None:
                raise RuntimeError("nvcc not found: cannot build csrc/libdrone2d.so (the CUDA extension is mandatory)")
            tmp = "%s.%d.tmp" % (LIB, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            if verbose:
                print(res.stderr)
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
