"""ctypes front for ``oracle/seeger_oracle.c`` -- the CPU parity checker.

TEST INFRASTRUCTURE ONLY (see the header of ``seeger_oracle.c``): the product
package ``cholupdates_b200`` never imports this module.

It exposes the same inner-kernel ABI as the reference's
``cholupdates.rank_1._seeger_impl_cython`` (``_seeger_impl_cython.pyx:22-25,
160-163``): ``update(L, v) -> None`` and ``downdate(L, v) -> None`` mutate both
arrays in place, do no argument validation beyond the memory-layout check
(``.pyx:68-77``) and raise ``numpy.linalg.LinAlgError`` when a downdate is not
positive definite (``.pyx:257-259``).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "seeger_oracle.c")
_LIB = os.path.join(HERE, "libseeger_oracle.so")
_lib = None


def _src_hash() -> str:
    import hashlib

    with open(_SRC, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _stale() -> bool:
    # contents, not file times: times do not survive the snapshot that carries the repo to the GPU box
    try:
        with open(_LIB + ".hash") as f:
            return not os.path.exists(_LIB) or f.read().strip() != _src_hash()
    except OSError:
        return True


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds).  Returns the library path.  Safe when several
    processes (torchrun ranks) call it at once: lock file, build into a private file, rename."""
    if not force and not _stale():
        return _LIB
    import fcntl

    with open(os.path.join(HERE, ".oracle_build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale():
                cc = os.environ.get("CC", "gcc")
                tmp = f"{_LIB}.tmp{os.getpid()}"
                # -ffp-contract=off: keep the restated BLAS loops un-fused, like the
                # reference-BLAS definitions they follow.
                try:
                    subprocess.check_call(
                        [cc, "-O2", "-fPIC", "-shared", "-ffp-contract=off", _SRC, "-o", tmp, "-lm"]
                    )
                    os.replace(tmp, _LIB)
                    with open(_LIB + ".hash", "w") as f:
                        f.write(_src_hash())
                finally:
                    if os.path.exists(tmp):
                        os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB)
        for suf in ("f64", "f32"):
            for op in ("update", "downdate"):
                fn = getattr(lib, f"seeger_oracle_{op}_{suf}")
                fn.restype = ctypes.c_int
                fn.argtypes = [
                    ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                    ctypes.c_void_p,
                ]
        _lib = lib
    return _lib


def _strides(L: np.ndarray):
    n = L.shape[0]
    if L.flags.f_contiguous:
        return 1, n
    if L.flags.c_contiguous:
        return n, 1
    raise ValueError(
        "Unsupported memory layout. L should either be Fortran- or C-contiguous."
    )


def _call(op: str, L: np.ndarray, v: np.ndarray) -> int:
    lib = _load()
    assert L.ndim == 2 and L.shape[0] == L.shape[1] and v.shape == (L.shape[0],)
    assert L.dtype == v.dtype and L.dtype in (np.float64, np.float32)
    assert v.flags.c_contiguous
    rs, cs = _strides(L)
    suf = "f64" if L.dtype == np.float64 else "f32"
    if L.shape[0] == 0:
        return 0
    return getattr(lib, f"seeger_oracle_{op}_{suf}")(
        L.ctypes.data, L.shape[0], rs, cs, v.ctypes.data
    )


def update(L: np.ndarray, v: np.ndarray) -> None:
    """In-place rank-1 update, restating ``_seeger_impl_cython.pyx:22-155``."""
    _call("update", L, v)


def downdate(L: np.ndarray, v: np.ndarray) -> None:
    """In-place rank-1 downdate, restating ``_seeger_impl_cython.pyx:160-330``."""
    if _call("downdate", L, v) != 0:
        raise np.linalg.LinAlgError("The downdated matrix is not positive definite.")
