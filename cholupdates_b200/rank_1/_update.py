"""Interface function for the symmetric rank-1 update (method dispatch).

Mirrors ``cholupdates.rank_1.update`` (reference ``src/cholupdates/rank_1/_update.py:10-185``).
"""

from __future__ import annotations

from ._seeger import update_seeger


def update(
    L,
    v,
    check_diag: bool = True,
    overwrite_L: bool = False,
    overwrite_v: bool = False,
    method: str = "seeger",
    **method_kwargs,
):
    r"""Update a Cholesky factorization after addition of a positive-semidefinite symmetric
    rank-1 matrix: given :math:`A = L L^T` and :math:`v`, compute :math:`L^+` with
    :math:`A + v v^T = L^+ (L^+)^T`.

    Drop-in for the reference's ``update`` (``_update.py:10-18``): same arguments, same return
    value, same exceptions.

    Parameters
    ----------
    L : (N, N) numpy.ndarray or CUDA torch.Tensor, float32 / float64, C- or F-contiguous
        Lower-triangular Cholesky factor.  The strict upper triangle is never read or written.
    v : (N,) array of the same kind and dtype
    check_diag : raise ``numpy.linalg.LinAlgError`` if ``L`` has zeros on its diagonal.
    overwrite_L, overwrite_v : allow the call to work in place (``_seeger.py:272-276``).
    method :
        ``"seeger"`` -- the algorithm of section 2 of [Seeger, 2008] (the reference's default);
        here it runs on the B200.  ``"b200"`` is the explicit name of the same GPU method.
        ``"cho_factor"`` -- the reference's :math:`O(N^3)` cross-check (``_update.py:157-169``):
        ``tril(L) tril(L)^T + v v^T`` re-factorised, here on the GPU (cuSOLVER through torch).
    method_kwargs : forwarded to ``update_seeger`` (``impl=None | "b200"``).

    Examples
    --------
    The reference's known-answer doctest (``_update.py:104-153``)::

        A = np.diag([1.0, 2.0, 3.0]) + 0.1
        L = np.linalg.cholesky(A)
        L_ud = cholupdates_b200.rank_1.update(L, np.array([1.0, 25.0, 10.0]))
        # [[ 1.44913767  0.          0.        ]
        #  [17.32064554 18.08577447  0.        ]
        #  [ 6.96966215  7.15374133  1.82969791]]
    """
    if method == "cho_factor":
        L_upd = _cho_factor_method(L, v, check_diag, **method_kwargs)
    elif method in ("seeger", "b200"):
        L_upd = update_seeger(
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
    """``method="cho_factor"`` (reference ``_update.py``: ``tril(L) tril(L)^T + v v^T`` re-factorised, strict
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
    A = Ltri @ Ltri.T + torch.outer(vt, vt)
    Lc, info = torch.linalg.cholesky_ex(A)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError(f"{int(info.item())}-th leading minor of the array is not positive definite")
    out = torch.tril(Lc) + torch.triu(Lt, 1)
    if as_numpy:
        return np.array(out.cpu().numpy(), order="F" if (L.flags.f_contiguous and not L.flags.c_contiguous) else "C")
    return out


update.available_methods = ["cho_factor", "seeger", "b200"]


def update_in_place(L, v, check_diag: bool = True, method: str = "seeger", **method_kwargs):
    """``update(..., overwrite_L=True, overwrite_v=True)``: the reference's "in-place" mode
    (``notebooks/benchmark_rank_1.ipynb:156-180``).  Returns ``L`` (the same object)."""
    return update(
        L, v, check_diag=check_diag, overwrite_L=True, overwrite_v=True, method=method,
        **method_kwargs,
    )
