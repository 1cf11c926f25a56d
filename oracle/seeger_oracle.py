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


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds).  Returns the library path."""
    if (
        force
        or not os.path.exists(_LIB)
        or os.path.getmtime(_LIB) < os.path.getmtime(_SRC)
    ):
        cc = os.environ.get("CC", "gcc")
        # -ffp-contract=off: keep the restated BLAS loops un-fused, like the
        # reference-BLAS definitions they follow.
        subprocess.check_call(
            [cc, "-O2", "-fPIC", "-shared", "-ffp-contract=off", _SRC, "-o", _LIB, "-lm"]
        )
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
