"""Build ``lib/libcholupdates_b200.so`` (the C-ABI library, hand-written sm_100a CUDA) in-tree.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the build container; the
resulting ``.so`` is git-ignored but travels to the GPU box with the repository snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libcholupdates_b200.so")
SOURCES = ["api.cu", "host.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _source_hash() -> str:
    """Content hash of everything the library is compiled from (file times do not survive the snapshot
    that carries the repository to the GPU box, contents do)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    deps.append(os.path.join(os.path.dirname(PKG), "include", "cholupdates_b200.h"))
    for d in deps:
        if os.path.isfile(d):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    try:
        with open(LIB + ".hash") as f:
            return f.read().strip() != _source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources.  Returns its path.

    Safe under ``torchrun``: the ranks of a job serialise on a lock file, the first one builds (into a
    file of its own, renamed into place), the others find the library fresh when they get the lock."""
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl

    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB  # another process built it while this one waited
            nvcc = _nvcc()
            if nvcc is None:
                if os.path.exists(LIB):
                    return LIB  # cannot rebuild here; use what travelled with the snapshot
                raise RuntimeError("nvcc not found and libcholupdates_b200.so is not built")
            tmp = f"{LIB}.tmp{os.getpid()}"
            cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
            if verbose:
                print(" ".join(cmd))
            try:
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)
                with open(LIB + ".hash", "w") as f:
                    f.write(_source_hash())
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force=True, verbose=True))
