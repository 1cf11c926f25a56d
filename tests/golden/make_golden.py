"""Generate the golden input/output vectors under ``tests/golden/`` by RUNNING THE
REFERENCE ITSELF (unmodified ``/root/reference/src/cholupdates`` + its Cython
kernels compiled by ``oracle/build_ref.py``).

Run here (the container that has ``/root/reference``); commit the outputs.  The
GPU box has no ``/root/reference`` -- tests there compare against these files.

    python tests/golden/make_golden.py

Inputs follow the reference's own test fixtures:
  * ``A`` / ``L``: ``tests/cholupdates/rank_1/conftest.py:27-51``
    (``utils.random_spd_eigendecomposition``, ``np.linalg.cholesky``)
  * update ``v``:  ``tests/cholupdates/rank_1/update/conftest.py:13-17``
  * downdate ``v``: ``tests/cholupdates/rank_1/downdate/conftest.py:14-19``
    (``utils.random_rank_1_downdate``)
Outputs are produced by ``cholupdates.rank_1.update/downdate(method="seeger",
impl="cython", overwrite_L=True, overwrite_v=True)`` for C- and F-ordered ``L``
and, for the fp64 cases, also by ``impl="python"``.

Only lower triangles are stored (packed with ``np.tril_indices``) to keep the
fixtures small.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_api  # noqa: E402

SEEDS = (0, 1, 2)
SIZES = (1, 2, 3, 5, 10, 100)
EXTRA = ((257, 0),)  # (N, seed): one size that is neither a power of two nor <= 256 (fp64 only)


def _case(ref, N, dtype, seed):
    rng = np.random.default_rng(seed)
    spectrum, Q = ref.utils.random_spd_eigendecomposition(N, dtype=dtype, rng=rng)
    A = Q @ np.diag(spectrum) @ Q.T
    L = np.linalg.cholesky(A)
    v_ud = rng.normal(scale=10, size=N).astype(dtype, copy=False)
    v_dd = ref.utils.random_rank_1_downdate(L, rng=rng)
    tri = np.tril_indices(N)
    out = {"L": L[tri], "v_ud": v_ud, "v_dd": v_dd}
    for order in ("C", "F"):
        for op, v in (("ud", v_ud), ("dd", v_dd)):
            fn = ref.rank_1.update if op == "ud" else ref.rank_1.downdate
            Lc, vc = np.array(L, order=order), v.copy()
            Lo = fn(Lc, vc, overwrite_L=True, overwrite_v=True, method="seeger", impl="cython")
            assert Lo is Lc
            out[f"L_{op}_{order}"] = Lo[tri]
            out[f"vout_{op}_{order}"] = vc
            if dtype == np.float64 and order == "F":
                Lp = fn(np.array(L, order=order), v.copy(), method="seeger", impl="python")
                out[f"L_{op}_python"] = Lp[tri]
    return out


def main():
    ref = ref_api.load_reference_api()
    assert ref is not None, "needs /root/reference"
    assert "cython" in ref.rank_1.update_seeger.available_impls
    blobs = {}
    for dtype, tag in ((np.float64, "f64"), (np.float32, "f32")):
        cases = [(N, s) for N in SIZES for s in SEEDS]
        if dtype == np.float64:
            cases += list(EXTRA)
        for N, seed in cases:
            for k, a in _case(ref, N, dtype, seed).items():
                blobs[f"{tag}/N{N}/seed{seed}/{k}"] = a
    path = os.path.join(HERE, "seeger_golden.npz")
    np.savez_compressed(path, **blobs)
    print(f"wrote {path}: {len(blobs)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")

    # The reference's two known-answer doctests (rank_1/_update.py:104-153,
    # rank_1/_downdate.py:107-148), re-run through the live reference so the
    # literals in tests/test_oracle.py are confirmed against it here.
    A = np.diag([1.0, 2.0, 3.0]) + 0.1
    import scipy.linalg

    L = scipy.linalg.cho_factor(A, lower=True)[0]
    L_ud = ref.rank_1.update(L, np.array([1.0, 25.0, 10.0]), method="seeger")
    print(np.array2string(np.tril(L_ud), precision=8))
    A = np.array([[7.77338976, 1.27256923, 1.58075291],
                  [1.27256923, 8.29126934, 0.80466256],
                  [1.58075291, 0.80466256, 13.65749896]])
    L = scipy.linalg.cholesky(A, lower=True)
    L_dd = ref.rank_1.downdate(L, np.array([1.60994441, 0.21482681, 0.78780241]), method="seeger")
    print(np.array2string(L_dd, precision=8))


if __name__ == "__main__":
    main()
