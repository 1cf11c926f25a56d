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
    if method in ("seeger", "b200"):
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


downdate.available_methods = ["seeger", "b200"]


def downdate_in_place(L, v, check_diag: bool = True, method: str = "seeger", **method_kwargs):
    """``downdate(..., overwrite_L=True, overwrite_v=True)``.  Returns ``L`` (the same object)."""
    return downdate(
        L, v, check_diag=check_diag, overwrite_L=True, overwrite_v=True, method=method,
        **method_kwargs,
    )
