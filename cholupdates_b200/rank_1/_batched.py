"""Batched rank-1 update / downdate of independent factors (GP / Kalman style).

New surface: the reference has no batched call, and a 3-D ``L`` must keep raising
``ValueError`` through ``update()`` (``tests/cholupdates/rank_1/update/test_update.py:88-97``),
so batches get their own entry points.  Each factor is processed by one CTA
(``csrc/small.cuh``); a failing factor is left untouched and reported in ``status``.
"""

from __future__ import annotations

import numpy as np

from .. import _lib
from ._arg_validation import is_torch_tensor
from ._seeger import _NOT_PD_MSG, _ZERO_DIAG_MSG


def _batched(op: str, L, v, check_diag, overwrite_L, overwrite_v, raise_on_error):
    import torch

    if not (is_torch_tensor(L) and is_torch_tensor(v)):
        raise TypeError("batched calls take torch CUDA tensors")
    if L.ndim != 3 or L.shape[1] != L.shape[2]:
        raise ValueError(f"`L` must have shape (B, N, N) (given shape: {tuple(L.shape)}).")
    if v.ndim != 2 or v.shape[0] != L.shape[0] or v.shape[1] != L.shape[1]:
        raise ValueError(
            f"`v` must have shape {(L.shape[0], L.shape[1])} but got {tuple(v.shape)}."
        )
    if L.dtype not in (torch.float32, torch.float64) or v.dtype != L.dtype:
        raise TypeError("`L` and `v` must both be float32 or both be float64.")
    if not (L.is_cuda and v.is_cuda and L.device == v.device):
        raise ValueError("`L` and `v` must live on the same CUDA device (no CPU fallback).")
    B, n = int(L.shape[0]), int(L.shape[1])
    if L.is_contiguous():
        layout = _lib.ROW_MAJOR
    elif L.transpose(1, 2).is_contiguous():
        layout = _lib.COL_MAJOR
    else:
        raise ValueError("each factor of `L` must be C- or F-contiguous and the batch dense.")
    if not v.is_contiguous():
        raise ValueError("`v` is not a contiguous array.")
    if not overwrite_L:
        L = L.clone(memory_format=torch.preserve_format)
    if not overwrite_v:
        v = v.clone()
    lib = _lib.require_device()
    suf = "f64" if L.dtype == torch.float64 else "f32"
    with torch.cuda.device(L.device):
        status = torch.empty(max(B, 1), dtype=torch.int32, device=L.device)
        rc = getattr(lib, f"chub_{op}_batched_{suf}")(
            L.data_ptr(), n, n, n * n, layout, v.data_ptr(), n, B,
            _lib.FLAG_CHECK_DIAG if check_diag else 0, status.data_ptr(),
            torch.cuda.current_stream(L.device).cuda_stream,
        )
        _lib.check(rc, f"chub_{op}_batched_{suf}")
    status = status[:B]
    if raise_on_error:
        bad = torch.nonzero(status).flatten()
        if bad.numel() > 0:
            b = int(bad[0].item())
            code = int(status[b].item())
            msg = _ZERO_DIAG_MSG if code == _lib.STATUS_ZERO_DIAG else _NOT_PD_MSG
            raise np.linalg.LinAlgError(f"{msg} (factor {b} of the batch; {bad.numel()} failed)")
        return L
    return L, status


def update_batched(L, v, check_diag=True, overwrite_L=False, overwrite_v=False,
                   raise_on_error=True):
    """Rank-1 update of ``B`` independent factors: ``L`` is ``(B, N, N)``, ``v`` is ``(B, N)``.

    Returns the updated ``L``; with ``raise_on_error=False`` returns ``(L, status)`` where
    ``status[b]`` is 0 (ok) or 1 (zero on the diagonal; factor ``b`` untouched).
    """
    return _batched("update", L, v, check_diag, overwrite_L, overwrite_v, raise_on_error)


def downdate_batched(L, v, check_diag=True, overwrite_L=False, overwrite_v=False,
                     raise_on_error=True):
    """Rank-1 downdate of ``B`` independent factors; ``status[b] == 2`` marks a factor whose
    downdate is not positive definite (left untouched)."""
    return _batched("downdate", L, v, check_diag, overwrite_L, overwrite_v, raise_on_error)
