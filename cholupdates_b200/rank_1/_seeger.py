"""Front-end of the Seeger rank-1 update / downdate on the B200.

Mirrors the reference's ``update_seeger`` / ``downdate_seeger``
(``src/cholupdates/rank_1/_seeger.py:37-296`` and ``:302-732``): the same signature, the same
checks in the same order with the same exception types and messages, the same copy-on-demand
semantics and return value.  What changes is the kernel behind it: instead of
``_seeger_impl_cython.update/downdate`` (``_seeger.py:17-34`` picks it as the default) the call
goes through the C ABI of ``include/cholupdates_b200.h`` to hand-written sm_100a kernels.

Inputs may be ``numpy.ndarray`` (staged to the GPU and back by ``chub_*_host_*``) or
``torch.Tensor`` on a CUDA device (zero-copy: ``data_ptr()`` on the current stream).
There is no CPU fallback.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from .. import _lib
from ._arg_validation import _has_zero_on_diagonal, _validate_update_args, is_torch_tensor

_update_available_impls = ["b200"]
_downdate_available_impls = ["b200"]

_ZERO_DIAG_MSG = "The given Cholesky factor `L` contains zeros on its diagonal"
_NOT_PD_MSG = "The downdated matrix is not positive definite."


# ----------------------------------------------------------------- helpers
def _dtype_name(x) -> str:
    if is_torch_tensor(x):
        return str(x.dtype).replace("torch.", "")
    return x.dtype.name


def _is_single_or_double(x) -> bool:
    if is_torch_tensor(x):
        import torch

        return x.dtype in (torch.float32, torch.float64)
    return x.dtype in (np.single, np.double)


def _layout_of(L) -> Optional[int]:
    """CHUB_ROW_MAJOR / CHUB_COL_MAJOR, or None if L is neither C- nor F-contiguous."""
    if is_torch_tensor(L):
        if L.is_contiguous():
            return _lib.ROW_MAJOR
        if L.t().is_contiguous():
            return _lib.COL_MAJOR
        return None
    if L.flags.f_contiguous and not L.flags.c_contiguous:
        return _lib.COL_MAJOR
    if L.flags.c_contiguous:
        return _lib.ROW_MAJOR
    return None


def _layout_and_ld(L, strided: bool):
    """(layout, leading dimension in elements), or None.  Dense C / F arrays always; with ``strided``
    also row- or column-major VIEWS whose inner stride is 1 and whose leading dimension exceeds N (a
    corner of a wider buffer): a superset of what the reference accepts (``_seeger.py:260-264``), which
    the C ABI has taken from the start (``ld`` argument)."""
    n = int(L.shape[0])
    lay = _layout_of(L)
    if lay is not None:
        return lay, max(n, 1)
    if not strided or n == 0:
        return None
    if is_torch_tensor(L):
        s0, s1 = int(L.stride(0)), int(L.stride(1))
    else:
        if L.strides[0] % L.itemsize or L.strides[1] % L.itemsize:
            return None
        s0, s1 = L.strides[0] // L.itemsize, L.strides[1] // L.itemsize
    if s1 == 1 and s0 >= n:
        return _lib.ROW_MAJOR, s0
    if s0 == 1 and s1 >= n:
        return _lib.COL_MAJOR, s1
    return None


def _from_dlpack(x):
    """Objects that are neither numpy arrays nor torch tensors but speak DLPack (CuPy arrays, JAX / numba
    device arrays ...) are viewed as torch CUDA tensors without a copy."""
    if isinstance(x, np.ndarray) or is_torch_tensor(x) or not hasattr(x, "__dlpack__"):
        return x
    import torch

    return torch.from_dlpack(x)


def _to_caller_type(result, like):
    """A freshly allocated result goes back in the caller's array type when that type can import DLPack."""
    if like is None or isinstance(like, np.ndarray) or is_torch_tensor(like):
        return result
    try:
        import importlib

        xp = importlib.import_module(type(like).__module__.partition(".")[0])
        return xp.from_dlpack(result)
    except Exception:  # noqa: BLE001
        return result


def _is_contiguous_vector(v) -> bool:
    if is_torch_tensor(v):
        return v.is_contiguous()
    return bool(v.flags.c_contiguous)


def _checks_after_diag(L, v, strided: bool = False) -> Optional[Exception]:
    """The reference's checks that come after the diagonal scan, first failure only
    (``_arg_validation.py:25-30`` then ``_seeger.py:243-269`` / ``:679-705``)."""
    if v.ndim != 1 or v.shape[0] != L.shape[0]:
        return ValueError(
            f"The shape of the given vector `v` is incompatible with the shape of the "
            f"given Cholesky factor `L`. Expected shape {(L.shape[0],)} but got "
            f"{tuple(v.shape)}."
        )
    if not _is_single_or_double(L):
        return TypeError(
            f"The given Cholesky factor `L` does not have dtype `np.single` or "
            f"`np.double` (given dtype: {_dtype_name(L)})"
        )
    if not _is_single_or_double(v):
        return TypeError(
            f"The given vector `v` does not have dtype `np.single` or `np.double` "
            f"(given dtype: {_dtype_name(v)})."
        )
    if L.dtype != v.dtype:
        return TypeError(
            f"`L` and `v` don't have the same dtype ({_dtype_name(L)} != {_dtype_name(v)})."
        )
    if _layout_and_ld(L, strided) is None:
        return ValueError(
            "`L` is neither Fortran- nor C-contiguous, i.e. the memory layout of the "
            "matrix is neither column- nor row-major."
        )
    if not _is_contiguous_vector(v):
        return ValueError("`v` is not a contiguous array.")
    return None


def _validate(L, v, check_diag: bool, strided: bool = False) -> bool:
    """Run the reference's checks in the reference's order.  Returns True if the diagonal scan
    is left to the kernel (CUDA tensors with nothing else wrong)."""
    if is_torch_tensor(L) != is_torch_tensor(v):
        raise TypeError("`L` and `v` must both be numpy arrays or both be torch tensors.")
    if is_torch_tensor(L):
        if not (L.is_cuda and v.is_cuda and L.device == v.device):
            raise ValueError(
                "torch inputs must live on the same CUDA device (the B200 path has no CPU "
                f"fallback); got L on {L.device}, v on {v.device}."
            )
        if L.ndim != 2 or L.shape[0] != L.shape[1]:
            _validate_update_args(L, v, check_diag=False)  # raises the reference's ValueError
        later = _checks_after_diag(L, v, strided)
        if later is not None:
            if check_diag and _has_zero_on_diagonal(L):
                raise np.linalg.LinAlgError(_ZERO_DIAG_MSG)
            raise later
        return bool(check_diag)
    # NumPy: literally the reference's sequence.
    _validate_update_args(L, v, check_diag)
    later = _checks_after_diag(L, v, strided)
    if later is not None:
        raise later
    return False


_ws_cache: dict = {}
_WS_CACHE_MAX = 32


def _torch_workspace(device, stream_id: int, nbytes: int):
    import torch

    key = (device.index, stream_id)
    buf = _ws_cache.pop(key, None)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
    _ws_cache[key] = buf  # (re-inserted last: the dict doubles as the LRU order)
    while len(_ws_cache) > _WS_CACHE_MAX:  # streams come and go: keep the most recently used scratch buffers only
        old_key = next(iter(_ws_cache))
        torch.cuda.synchronize(old_key[0])  # (rare path: a kernel on that stream may still be using the buffer)
        _ws_cache.pop(old_key)
    return buf


def release_workspaces() -> None:
    """Drop the cached per-(device, stream) scratch buffers (they are re-created on demand)."""
    _ws_cache.clear()


def _raise_for_status(status: int, op: str) -> None:
    if status == _lib.STATUS_OK:
        return
    if status == _lib.STATUS_ZERO_DIAG:
        raise np.linalg.LinAlgError(_ZERO_DIAG_MSG)
    if status == _lib.STATUS_NOT_PD:
        raise np.linalg.LinAlgError(_NOT_PD_MSG)
    raise RuntimeError(f"cholupdates_b200: {op} kernel reported internal status {status}")


def _run(op: str, L, v, diag_on_device: bool, strided: bool = False, L_out=None) -> None:
    """In-place kernel call: the replacement for ``_update_impl_default(L, v)``
    (``_seeger.py:279-294``) / ``_downdate_impl_default(L, v)`` (``:715-730``)."""
    lib = _lib.require_device()
    n = int(L.shape[0])
    layout, ld = _layout_and_ld(L, strided)
    flags = _lib.FLAG_CHECK_DIAG if diag_on_device else 0
    if is_torch_tensor(L):
        import torch

        suf = "f64" if L.dtype == torch.float64 else "f32"
        with torch.cuda.device(L.device):
            stream = torch.cuda.current_stream(L.device)
            nbytes = lib.chub_workspace_bytes(1 if suf == "f64" else 0, n)
            ws = _torch_workspace(L.device, stream.cuda_stream, nbytes)
            status = torch.empty(1, dtype=torch.int32, device=L.device)
            rc = getattr(lib, f"chub_{op}_{suf}")(
                L.data_ptr(), n, ld, layout, v.data_ptr(), flags, ws.data_ptr(), ws.numel(),
                status.data_ptr(), stream.cuda_stream,
            )
            _lib.check(rc, f"chub_{op}_{suf}")
            code = int(status.item())  # one 4-byte D2H read; L is untouched on any non-zero code
    else:
        suf = "f64" if L.dtype == np.double else "f32"
        st = ctypes.c_int32(0)
        if L_out is None:
            rc = getattr(lib, f"chub_{op}_host_{suf}")(
                L.ctypes.data, n, ld, layout, v.ctypes.data, flags, ctypes.byref(st)
            )
        else:  # copy mode: read L, write L_out (dense, same order)
            rc = getattr(lib, f"chub_{op}_host_copy_{suf}")(
                L.ctypes.data, ld, L_out.ctypes.data, max(n, 1), n, layout, v.ctypes.data, flags,
                ctypes.byref(st)
            )
        _lib.check(rc, f"chub_{op}_host_{suf}")
        code = int(st.value)
    _raise_for_status(code, op)


def _copy_L(L):
    if is_torch_tensor(L):
        import torch

        return L.clone(memory_format=torch.preserve_format)  # == copy(order="K")
    return L.copy(order="K")


def _copy_v(v):
    return v.clone() if is_torch_tensor(v) else v.copy()


def _dispatch(op: str, L, v, check_diag, overwrite_L, overwrite_v, impl, available, strided=False):
    L_given = L
    L, v = _from_dlpack(L), _from_dlpack(v)  # CuPy & co: zero-copy views as torch CUDA tensors
    diag_on_device = _validate(L, v, check_diag, strided)

    # Copy on demand  (_seeger.py:272-276 / :708-712).  NumPy factors are not copied on the host first: the
    # host entry point reads L and delivers the result into a fresh array of the same memory order
    # (chub_*_host_copy_*; only the strict upper triangle is copied host-side, overlapped with the transfers).
    L_out = None
    if not overwrite_L:
        if isinstance(L, np.ndarray) and (impl is None or impl == "b200") and L.shape[0] > 0:
            layout, _ = _layout_and_ld(L, strided)
            L_out = np.empty(L.shape, dtype=L.dtype, order="F" if layout == _lib.COL_MAJOR else "C")
        else:
            L = _copy_L(L)
    if not overwrite_v:
        v = _copy_v(v)

    if impl is None or impl == "b200":
        _run(op, L, v, diag_on_device, strided, L_out)
        if L_out is not None:
            L = L_out
    elif impl in ("cython", "python"):
        raise NotImplementedError(
            f"The {impl.capitalize()} implementation is not available in cholupdates_b200 "
            "(GPU only, no CPU fallback)."
        )
    else:
        raise NotImplementedError(
            f"Unknown implementation '{impl}'. Available implementations: "
            f"{', '.join(available)}"
        )
    if L_given is not L and not isinstance(L_given, np.ndarray) and not is_torch_tensor(L_given):
        return L_given if overwrite_L else _to_caller_type(L, L_given)  # DLPack callers get their own type back
    return L


# ------------------------------------------------------------------ public
def update_seeger(
    L,
    v,
    check_diag: bool = True,
    overwrite_L: bool = False,
    overwrite_v: bool = False,
    impl: Optional[str] = None,
    strided: bool = False,
):
    r"""Rank-1 update :math:`L^+ (L^+)^T = L L^T + v v^T` in :math:`O(N^2)` by the algorithm of
    section 2 of [Seeger, 2008], on the B200.

    Same contract as the reference's ``update_seeger`` (``_seeger.py:37-296``):

    * ``L``: ``(N, N)``, ``single`` or ``double``, C- or F-contiguous; only the lower triangle
      is read or written, the strict upper triangle passes through bit-identical.
    * ``v``: ``(N,)``, same dtype, contiguous; used as scratch if ``overwrite_v``.
    * ``check_diag``: raise ``numpy.linalg.LinAlgError`` if ``L`` has a zero on its diagonal.
    * returns the updated factor: ``L`` itself if ``overwrite_L`` else a copy that keeps
      ``L``'s memory order.
    * ``impl``: ``None`` or ``"b200"``.
    * ``strided`` (new, default off so that the reference's ``ValueError`` for non-contiguous input is
      kept): also accept a row- or column-major VIEW with unit inner stride and a leading dimension
      larger than ``N`` (e.g. ``big[:N, :N]``); nothing outside the ``N x N`` lower triangle is touched.
    * Besides ``numpy.ndarray`` and ``torch.Tensor``, any CUDA array that speaks DLPack (CuPy ...) is
      accepted zero-copy; a non-overwriting call returns the result in the caller's array type.

    Raises ``ValueError`` / ``TypeError`` / ``numpy.linalg.LinAlgError`` /
    ``NotImplementedError`` exactly where the reference does (``_seeger.py:241-294``).
    """
    return _dispatch(
        "update", L, v, check_diag, overwrite_L, overwrite_v, impl, _update_available_impls, strided
    )


update_seeger.available_impls = _update_available_impls


def downdate_seeger(
    L,
    v,
    check_diag: bool = True,
    overwrite_L: bool = False,
    overwrite_v: bool = False,
    impl: Optional[str] = None,
    strided: bool = False,
):
    r"""Rank-1 downdate :math:`L^- (L^-)^T = L L^T - v v^T` in :math:`O(N^2)` by the algorithm of
    section 3 of [Seeger, 2008], on the B200.

    Same contract as the reference's ``downdate_seeger`` (``_seeger.py:302-732``); additionally
    raises ``numpy.linalg.LinAlgError("The downdated matrix is not positive definite.")`` when
    :math:`\rho^2 = 1 - \|L^{-1} v\|^2 \le 0` (``_seeger_impl_cython.pyx:255-259``), in which case
    ``L`` has not been modified.
    """
    return _dispatch(
        "downdate", L, v, check_diag, overwrite_L, overwrite_v, impl, _downdate_available_impls, strided
    )


downdate_seeger.available_impls = _downdate_available_impls
