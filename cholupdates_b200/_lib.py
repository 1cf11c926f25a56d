"""ctypes binding of the C ABI in ``include/cholupdates_b200.h``.

There is NO CPU fallback: if the library cannot be loaded, or no CUDA device is visible, the
calls raise ``RuntimeError``.
"""

from __future__ import annotations

import ctypes
import os
import threading

from . import _build

ROW_MAJOR = 0
COL_MAJOR = 1
FLAG_CHECK_DIAG = 1
STATUS_OK, STATUS_ZERO_DIAG, STATUS_NOT_PD, STATUS_INTERNAL = 0, 1, 2, 3

_lock = threading.Lock()
_lib = None

_vp, _i64, _int, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
_SIG_SINGLE = [_vp, _i64, _i64, _int, _vp, _int, _vp, _sz, _vp, _vp]
_SIG_BATCHED = [_vp, _i64, _i64, _i64, _int, _vp, _i64, _i64, _int, _vp, _vp]
_SIG_HOST = [_vp, _i64, _i64, _int, _vp, _int, ctypes.POINTER(ctypes.c_int32)]
_SIG_HOST_COPY = [_vp, _i64, _vp, _i64, _i64, _int, _vp, _int, ctypes.POINTER(ctypes.c_int32)]


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer and nvcc is present) the C-ABI library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        try:
            path = os.environ.get("CHUB_LIB") or _build.build()  # CHUB_LIB: A/B-test another build
        except Exception as exc:  # noqa: BLE001
            raise RuntimeError(
                "cholupdates_b200: the CUDA library libcholupdates_b200.so is not available "
                f"({exc}); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback."
            ) from exc
        lib = ctypes.CDLL(path)
        lib.chub_version.restype = _int
        lib.chub_last_error.restype = ctypes.c_char_p
        lib.chub_device_count.restype = _int
        lib.chub_device_info.restype = _int
        lib.chub_device_info.argtypes = [ctypes.POINTER(_int), ctypes.POINTER(_i64),
                                         ctypes.POINTER(_int), ctypes.POINTER(_int)]
        lib.chub_workspace_bytes.restype = _sz
        lib.chub_workspace_bytes.argtypes = [_int, _i64]
        lib.chub_set_path.restype = _int
        lib.chub_set_path.argtypes = [_int]
        lib.chub_launch_count.restype = _i64
        lib.chub_launch_count.argtypes = [_int]
        lib.chub_host_release.restype = _int
        for op in ("update", "downdate"):
            for suf in ("f64", "f32"):
                f = getattr(lib, f"chub_{op}_{suf}")
                f.restype, f.argtypes = _int, _SIG_SINGLE
                f = getattr(lib, f"chub_{op}_batched_{suf}")
                f.restype, f.argtypes = _int, _SIG_BATCHED
                f = getattr(lib, f"chub_{op}_host_{suf}")
                f.restype, f.argtypes = _int, _SIG_HOST
                f = getattr(lib, f"chub_{op}_host_copy_{suf}", None)  # (absent in older A/B builds, CHUB_LIB)
                if f is not None:
                    f.restype, f.argtypes = _int, _SIG_HOST_COPY
        _lib = lib
        return lib


def require_device() -> ctypes.CDLL:
    lib = load()
    if lib.chub_device_count() <= 0:
        raise RuntimeError(
            "cholupdates_b200: no CUDA device is visible; the B200 path has no CPU fallback."
        )
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().chub_last_error().decode(errors="replace")
        raise RuntimeError(f"cholupdates_b200: {what} failed (CUDA error {rc}): {msg}")
