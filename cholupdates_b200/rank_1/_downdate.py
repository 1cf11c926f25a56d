"""Interface function for the symmetric rank-1 downdate (method dispatch).

Mirrors ``cholupdates.rank_1.downdate`` (reference ``src/cholupdates/rank_1/_downdate.py:12-178``).
"""

from __future__ import annotations

from ._seeger import downdate_seeger


def downdate(
    L,
    v,
    check_diag: bool = True,
    overwrite_L: bool = False,
    overwrite_v: bool = False,
    method: str = "seeger",
    **method_kwargs,
):
    r"""Update a Cholesky factorization after subtraction of a positive-semidefinite symmetric
    rank-1 matrix: given :math:`A = L L^T` and :math:`v`, compute :math:`L^-` with
    :math:`A - v v^T = L^- (L^-)^T`.

    Drop-in for the reference's ``downdate`` (``_downdate.py:12-20``).  Raises
    ``numpy.linalg.LinAlgError`` if the downdated matrix is not positive definite
    (``_seeger_impl_cython.pyx:255-259``); ``L`` is then unmodified.  See
    :func:`cholupdates_b200.rank_1.update` for the arguments.

    Examples
    --------
    The reference's known-answer doctest (``_downdate.py:107-148``)::

        A = np.array([[ 7.77338976,  1.27256923,  1.58075291],
                      [ 1.27256923,  8.29126934,  0.80466256],
                      [ 1.58075291,  0.80466256, 13.65749896]])
        v = np.array([1.60994441, 0.21482681, 0.78780241])
        L_dd = cholupdates_b200.rank_1.downdate(np.linalg.cholesky(A), v)
        # [[2.27628398 0.         0.        ]
        #  [0.40711529 2.8424243  0.        ]
        #  [0.13725652 0.20389013 3.6022848 ]]
    """
    if method == "cho_factor":
        L_upd = _cho_factor_method(L, v, check_diag, **method_kwargs)
    elif method in ("seeger", "b200"):
        L_upd = downdate_seeger(
            L,
            v,
            check_diag=check_diag,
            overwrite_L=overwrite_L,
            overwrite_v=overwrite_v,
            **method_kwargs,
        )
    else:
        raise NotImplementedError(f"Unknown method: '{method}'")

    return L_upd


def _cho_factor_method(L, v, check_diag, **method_kwargs):
    """``method="cho_factor"`` (reference ``_downdate.py``: ``tril(L) tril(L)^T - v v^T`` re-factorised, strict
    upper triangle of ``L`` restored): the O(N^3) cross-check, on the GPU (cuSOLVER through torch -- a
    library call is what this convenience path is in the reference, too).  Ignores ``overwrite_*`` like
    the reference.  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
    import numpy as np
    import torch

    from .. import _lib
    from ._arg_validation import _validate_update_args, is_torch_tensor

    _validate_update_args(L, v, check_diag)
    if method_kwargs:
        raise TypeError(f"unexpected keyword arguments for method 'cho_factor': {sorted(method_kwargs)}")
    _lib.require_device()
    as_numpy = not is_torch_tensor(L)
    Lt = torch.as_tensor(np.ascontiguousarray(L)).cuda() if as_numpy else L
    vt = torch.as_tensor(np.ascontiguousarray(v)).cuda() if not is_torch_tensor(v) else v
    Ltri = torch.tril(Lt)
    A = Ltri @ Ltri.T - torch.outer(vt, vt)
    Lc, info = torch.linalg.cholesky_ex(A)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError(f"{int(info.item())}-th leading minor of the array is not positive definite")
    out = torch.tril(Lc) + torch.triu(Lt, 1)
    if as_numpy:
        return np.array(out.cpu().numpy(), order="F" if (L.flags.f_contiguous and not L.flags.c_contiguous) else "C")
    return out


downdate.available_methods = ["cho_factor", "seeger", "b200"]


def downdate_in_place(L, v, check_diag: bool = True, method: str = "seeger", **method_kwargs):
    """``downdate(..., overwrite_L=True, overwrite_v=True)``.  Returns ``L`` (the same object)."""
    return downdate(
        L, v, check_diag=check_diag, overwrite_L=True, overwrite_v=True, method=method,
        **method_kwargs,
    )
