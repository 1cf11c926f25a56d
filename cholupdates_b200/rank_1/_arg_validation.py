"""Argument validation shared by the rank-1 entry points.

Mirrors the reference's ``_validate_update_args`` (``src/cholupdates/rank_1/
_arg_validation.py:6-30``): same order, same exception types, same messages.  Works on
``numpy.ndarray`` and on ``torch.Tensor`` (both expose ``ndim`` / ``shape``).
"""

from __future__ import annotations

import numpy as np


def is_torch_tensor(x) -> bool:
    return type(x).__module__.partition(".")[0] == "torch" and hasattr(x, "data_ptr")


def _has_zero_on_diagonal(L) -> bool:
    if is_torch_tensor(L):
        return bool((L.diagonal() == 0).any().item())
    return bool(np.any(np.diag(L) == 0))


def _validate_update_args(L, v, check_diag: bool, diag_checked_on_device: bool = False):
    # Validate L  (_arg_validation.py:12-16)
    if L.ndim != 2 or L.shape[0] != L.shape[1]:
        raise ValueError(
            f"The given Cholesky factor `L` is not a square matrix (given shape: "
            f"{tuple(L.shape)})."
        )

    # (_arg_validation.py:18-22)  For CUDA tensors the scan is fused into the kernel
    # (status word CHUB_STATUS_ZERO_DIAG) unless a later check would fail as well, in which
    # case it is done here so that the reference's order of exceptions is kept.
    if check_diag and not diag_checked_on_device:
        if _has_zero_on_diagonal(L):
            raise np.linalg.LinAlgError(
                "The given Cholesky factor `L` contains zeros on its diagonal"
            )

    # Validate v  (_arg_validation.py:25-30)
    if v.ndim != 1 or v.shape[0] != L.shape[0]:
        raise ValueError(
            f"The shape of the given vector `v` is incompatible with the shape of the "
            f"given Cholesky factor `L`. Expected shape {(L.shape[0],)} but got "
            f"{tuple(v.shape)}."
        )
