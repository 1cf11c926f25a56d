"""Pin the oracle (``oracle/seeger_oracle.c``) before anything is compared with it.

CPU-only.  Three anchors:
  1. the reference's two 3x3 doctest known-answer vectors,
  2. golden vectors produced by the reference itself (tests/golden/make_golden.py),
  3. the reference's compiled Cython kernels (oracle/_ref) run live on seeded inputs.
Tolerance: relative Frobenius error over the lower triangle <= 1e-12 * sqrt(n) for
fp64 (the north-star tolerance); 1e-4 * sqrt(n)-free fixed 2e-5 for fp32.
"""

import numpy as np
import pytest

import build_ref
import seeger_oracle as oracle
from helpers import rel_fro, unpack_tril

TOL64 = 1e-12


def tol(dtype, n):
    return TOL64 * np.sqrt(max(n, 1)) if dtype == np.float64 else 2e-5


# ---------------------------------------------------------------- 1. doctests
def test_doctest_update_known_answer():
    # reference: src/cholupdates/rank_1/_update.py:104-153
    A = np.diag([1.0, 2.0, 3.0]) + 0.1
    L = np.linalg.cholesky(A)
    v = np.array([1.0, 25.0, 10.0])
    for order in "CF":
        Lc, vc = np.array(L, order=order), v.copy()
        oracle.update(Lc, vc)
        expect = np.array(
            [[1.44913767, 0.0, 0.0], [17.32064554, 18.08577447, 0.0], [6.96966215, 7.15374133, 1.82969791]]
        )
        np.testing.assert_allclose(np.tril(Lc), expect, rtol=0, atol=5.1e-9)  # doctest prints 8 decimals
        np.testing.assert_allclose(Lc @ Lc.T, A + np.outer(v, v), rtol=1e-12)


def test_doctest_downdate_known_answer():
    # reference: src/cholupdates/rank_1/_downdate.py:107-148
    A = np.array(
        [[7.77338976, 1.27256923, 1.58075291], [1.27256923, 8.29126934, 0.80466256], [1.58075291, 0.80466256, 13.65749896]]
    )
    v = np.array([1.60994441, 0.21482681, 0.78780241])
    L = np.linalg.cholesky(A)
    for order in "CF":
        Lc, vc = np.array(L, order=order), v.copy()
        oracle.downdate(Lc, vc)
        expect = np.array(
            [[2.27628398, 0.0, 0.0], [0.40711529, 2.8424243, 0.0], [0.13725652, 0.20389013, 3.6022848]]
        )
        np.testing.assert_allclose(Lc, expect, rtol=0, atol=5.1e-9)  # doctest prints 8 decimals


# ------------------------------------------------------------ 2. golden files
def _golden_cases(golden):
    seen = set()
    for key in golden.files:
        tag, N, seed, _ = key.split("/")
        seen.add((tag, int(N[1:]), int(seed[4:])))
    return sorted(seen)


def test_golden_covers_reference_grid(golden):
    cases = _golden_cases(golden)
    for tag in ("f64", "f32"):
        for N in (1, 2, 3, 5, 10, 100):
            for seed in (0, 1, 2):
                assert (tag, N, seed) in cases
    assert ("f64", 257, 0) in cases


@pytest.mark.parametrize("op", ["ud", "dd"])
@pytest.mark.parametrize("order", ["C", "F"])
def test_oracle_matches_golden(golden, op, order):
    for tag, N, seed in _golden_cases(golden):
        dtype = np.float64 if tag == "f64" else np.float32
        pre = f"{tag}/N{N}/seed{seed}/"
        L = unpack_tril(golden[pre + "L"], N, order)
        v = golden[pre + f"v_{op}"].copy()
        (oracle.update if op == "ud" else oracle.downdate)(L, v)
        L_ref = unpack_tril(golden[pre + f"L_{op}_{order}"], N)
        assert rel_fro(L, L_ref) <= tol(dtype, N), (tag, N, seed)
        # v on return is scratch; the Cython path leaves rotg's z_k there
        v_ref = golden[pre + f"vout_{op}_{order}"]
        denom = max(np.linalg.norm(v_ref.astype(np.float64)), 1e-300)
        assert np.linalg.norm(v.astype(np.float64) - v_ref) / denom <= 50 * tol(dtype, N), (tag, N, seed)
        assert np.all(np.triu(L, 1) == 0)  # strict upper triangle never written


def test_golden_cython_vs_python_impl_agree(golden):
    # sanity on the fixtures themselves: the reference's two impls agree to tolerance
    for tag, N, seed in _golden_cases(golden):
        if tag != "f64":
            continue
        pre = f"{tag}/N{N}/seed{seed}/"
        for op in ("ud", "dd"):
            a = unpack_tril(golden[pre + f"L_{op}_F"], N)
            b = unpack_tril(golden[pre + f"L_{op}_python"], N)
            assert rel_fro(a, b) <= tol(np.float64, N)


# -------------------------------------------------- 3. live reference kernels
ref = build_ref.load()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


def _factor(n, seed, dtype=np.float64):
    g = np.random.default_rng(seed)
    L = np.tril(g.standard_normal((n, n)) / np.sqrt(n), -1)
    L[np.diag_indices(n)] = 1.0 + g.random(n)
    return L.astype(dtype)


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 7, 33, 300, 1000])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_matches_live_reference_update(n, order, dtype):
    L0 = _factor(n, n, dtype)
    v0 = np.random.default_rng(n + 1).standard_normal(n).astype(dtype)
    La, va = np.array(L0, order=order), v0.copy()
    Lb, vb = np.array(L0, order=order), v0.copy()
    oracle.update(La, va)
    ref.update(Lb, vb)
    assert rel_fro(La, Lb) <= tol(dtype, n)
    assert np.linalg.norm(va.astype(float) - vb) <= 50 * tol(dtype, n) * max(np.linalg.norm(vb.astype(float)), 1)


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 7, 33, 300, 1000])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_matches_live_reference_downdate(n, order, dtype):
    L0 = _factor(n, n, dtype)
    u = np.random.default_rng(n + 2).standard_normal(n)
    u *= 0.8 / np.linalg.norm(u)
    v0 = (L0.astype(np.float64) @ u).astype(dtype)
    La, va = np.array(L0, order=order), v0.copy()
    Lb, vb = np.array(L0, order=order), v0.copy()
    oracle.downdate(La, va)
    ref.downdate(Lb, vb)
    assert rel_fro(La, Lb) <= tol(dtype, n)


@needs_ref
@pytest.mark.parametrize("order", ["C", "F"])
def test_not_positive_definite_agrees_with_reference(order):
    n = 50
    L0 = _factor(n, 3)
    u = np.random.default_rng(5).standard_normal(n)
    u *= 1.2 / np.linalg.norm(u)  # ||L^-1 v|| = 1.2, mirrors tests/.../test_downdate.py:155-175
    v0 = L0 @ u
    for impl in (oracle, ref):
        L, v = np.array(L0, order=order), v0.copy()
        with pytest.raises(np.linalg.LinAlgError):
            impl.downdate(L, v)
        np.testing.assert_array_equal(L, np.array(L0, order=order))  # L untouched
        np.testing.assert_allclose(v, u, rtol=1e-10)  # v now holds p


def test_upper_triangle_never_touched():
    # mirrors tests/cholupdates/rank_1/update/test_update.py:25-55
    n = 20
    L0 = _factor(n, 9)
    v0 = np.random.default_rng(10).standard_normal(n)
    junk = np.triu(np.random.default_rng(11).standard_normal((n, n)), 1)
    for op in (oracle.update, oracle.downdate):
        vv = v0 if op is oracle.update else 0.3 * (L0 @ (v0 / np.linalg.norm(v0)))
        for order in "CF":
            La, Lb = np.array(L0, order=order), np.array(L0 + junk, order=order)
            op(La, vv.copy())
            op(Lb, vv.copy())
            np.testing.assert_array_equal(np.triu(Lb, 1), junk)
            np.testing.assert_array_equal(np.tril(La), np.tril(Lb))


# -------------------------------------------------- 4. the config-1 input generator (restated)
@pytest.mark.skipif(not build_ref.reference_present(), reason="/root/reference is not present")
@pytest.mark.parametrize("n", [50, 1000])
def test_config1_inputs_match_reference_generator(n):
    """helpers.config1_inputs / bench.config1_inputs restate the reference's
    random_spd_matrix(fast=True) (src/cholupdates/utils.py:78-90) because the reference package does not
    travel to the GPU box: the restatement must reproduce the live generator bit for bit."""
    import scipy.linalg

    import ref_api
    from helpers import config1_inputs

    api = ref_api.load_reference_api()
    A = api.utils.random_spd_matrix(n, fast=True, rng=np.random.default_rng(0))
    L_ref = scipy.linalg.cholesky(A, lower=True)
    L, v = config1_inputs(n)
    np.testing.assert_array_equal(L, L_ref)
    np.testing.assert_array_equal(v, np.random.default_rng(1).normal(scale=10.0, size=n))
    import bench

    Lb, vb = bench.config1_inputs(n)
    np.testing.assert_array_equal(Lb, L)
    np.testing.assert_array_equal(vb, v)
