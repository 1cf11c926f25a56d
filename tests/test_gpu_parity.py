"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped front-end)
against the oracle on identical seeded inputs, against the committed golden vectors, and through
the reference's own property tests (tests/cholupdates/rank_1/{update,downdate}/test_*.py).

Tolerances are those of helpers.tol_for: relative Frobenius error over the lower triangle
<= 1e-12 * sqrt(n) (fp64, the north-star bar) / 1e-5 * sqrt(n) (fp32).
"""

import numpy as np
import pytest

import seeger_oracle as oracle
from helpers import (config1_inputs, downdate_vector, rel_fro, rel_vec, spd_factor, synth_factor, tol_for,
                     unpack_tril)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import cholupdates_b200  # noqa: E402
from cholupdates_b200 import _lib  # noqa: E402

rank_1 = cholupdates_b200.rank_1
DTYPES = [np.float64, np.float32]


def _oracle(op, L, v):
    Lr, vr = np.array(L, order="F"), v.copy()
    getattr(oracle, op)(Lr, vr)
    return Lr, vr


def _run(op, L, v, kind, **kw):
    """kind: 'numpy' (host staging path) or 'torch' (device pointers)."""
    fn = getattr(rank_1, op)
    if kind == "numpy":
        return fn(L, v, **kw)
    Lt, vt = torch.from_numpy(np.ascontiguousarray(L)).cuda(), torch.from_numpy(v.copy()).cuda()
    if L.flags.f_contiguous and not L.flags.c_contiguous:
        Lt = torch.from_numpy(np.ascontiguousarray(L.T)).cuda().t()  # column-major CUDA tensor
    out = fn(Lt, vt, **kw)
    return out.cpu().numpy()


# ------------------------------------------------- reference grid (N, dtype, order, seed)
@pytest.mark.parametrize("N", [1, 2, 3, 5, 10, 100])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_reference_grid_update(N, dtype, order, kind):
    for seed in range(3):
        A, L = spd_factor(N, seed, dtype)
        v = np.random.default_rng(seed + 100).normal(scale=10, size=N).astype(dtype)
        L_in = np.array(L, order=order)
        L_ud = _run("update", L_in, v, kind)
        L_ref, _ = _oracle("update", L, v)
        assert rel_fro(L_ud, L_ref) <= tol_for(dtype, N)
        # test_update.py:11-22  valid square root, positive diagonal
        rtol = 1e-7 if dtype == np.float64 else 1e-4
        Lt = np.tril(L_ud).astype(np.float64)
        Aud = A.astype(np.float64) + np.outer(v, v).astype(np.float64)
        assert np.linalg.norm(Lt @ Lt.T - Aud) <= 10 * rtol * np.linalg.norm(Aud)
        assert np.all(np.diag(L_ud) > 0)
        # test_update_seeger.py:9-19  memory order is kept
        if kind == "numpy" and N > 1:
            assert L_ud.flags.c_contiguous == L_in.flags.c_contiguous
            assert L_ud.flags.f_contiguous == L_in.flags.f_contiguous


@pytest.mark.parametrize("N", [1, 2, 3, 5, 10, 100])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_reference_grid_downdate(N, dtype, order, kind):
    for seed in range(3):
        A, L = spd_factor(N, seed, dtype)
        v = downdate_vector(L, seed + 200)
        L_dd = _run("downdate", np.array(L, order=order), v, kind)
        L_ref, _ = _oracle("downdate", L, v)
        assert rel_fro(L_dd, L_ref) <= tol_for(dtype, N) * (1 if dtype == np.float64 else 30)
        rtol = 1e-7 if dtype == np.float64 else 3e-2
        Lt = np.tril(L_dd).astype(np.float64)
        Add = A.astype(np.float64) - np.outer(v, v).astype(np.float64)
        assert np.linalg.norm(Lt @ Lt.T - Add) <= 10 * rtol * np.linalg.norm(Add)
        assert np.all(np.diag(L_dd) > 0)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("op", ["ud", "dd"])
@pytest.mark.parametrize("order", ["C", "F"])
def test_matches_golden_vectors(golden, op, order):
    seen = sorted({tuple(k.split("/")[:3]) for k in golden.files})
    for tag, Ns, seeds in seen:
        N, dtype = int(Ns[1:]), (np.float64 if tag == "f64" else np.float32)
        pre = f"{tag}/{Ns}/{seeds}/"
        L = unpack_tril(golden[pre + "L"], N, order)
        v = golden[pre + f"v_{op}"].copy()
        fn = rank_1.update if op == "ud" else rank_1.downdate
        out = fn(L, v, overwrite_L=True, overwrite_v=True)
        assert out is L
        L_ref = unpack_tril(golden[pre + f"L_{op}_{order}"], N)
        slack = 1 if dtype == np.float64 else 30
        assert rel_fro(L, L_ref) <= tol_for(dtype, N) * slack, (tag, N, seeds)
        assert np.all(np.triu(L, 1) == 0)
        if dtype == np.float64:  # v on return: rotg's z_k, as the Cython path leaves it
            v_ref = golden[pre + f"vout_{op}_{order}"]
            assert rel_vec(v, v_ref) <= tol_for(dtype, N), (tag, N, seeds, rel_vec(v, v_ref))


# ------------------------------------------------- sizes across kernel paths, both layouts
SIZES = [31, 32, 33, 64, 100, 257, 511, 512, 513, 514, 600, 1000, 1022, 1025, 2049, 3000]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_update_sizes(n, order, dtype):
    L = synth_factor(n, n, dtype)
    v = np.random.default_rng(n + 1).standard_normal(n).astype(dtype)
    L_ref, v_ref = _oracle("update", L, v)
    for kind in ("numpy", "torch"):
        got = _run("update", np.array(L, order=order), v, kind)
        assert rel_fro(got, L_ref) <= tol_for(dtype, n), (kind,)
        np.testing.assert_array_equal(np.triu(got, 1), 0)
    if dtype == np.float64:
        vv = v.copy()
        rank_1.update(np.array(L, order=order), vv, overwrite_v=True)
        assert rel_vec(vv, v_ref) <= tol_for(dtype, n), rel_vec(vv, v_ref)  # v: the same bar as L


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_downdate_sizes(n, order, dtype):
    L = synth_factor(n, n, dtype)
    v = downdate_vector(L, n + 2)
    L_ref, v_ref = _oracle("downdate", L, v)
    slack = 1 if dtype == np.float64 else 30
    for kind in ("numpy", "torch"):
        got = _run("downdate", np.array(L, order=order), v, kind)
        assert rel_fro(got, L_ref) <= tol_for(dtype, n) * slack, (kind,)
        np.testing.assert_array_equal(np.triu(got, 1), 0)
    if dtype == np.float64:
        vv = v.copy()
        rank_1.downdate(np.array(L, order=order), vv, overwrite_v=True)
        assert rel_vec(vv, v_ref) <= tol_for(dtype, n), rel_vec(vv, v_ref)


@pytest.mark.parametrize("path", [1, 2, 3])
@pytest.mark.parametrize("n", [300, 777, 1536])
@pytest.mark.parametrize("order", ["C", "F"])
def test_forced_kernel_paths_agree(path, n, order):
    """1 = single-CTA kernels, 2 = launch-per-panel kernels, 3 = persistent dataflow kernel."""
    lib = _lib.load()
    L = synth_factor(n, 5)
    v = np.random.default_rng(6).standard_normal(n)
    vd = downdate_vector(L, 7)
    prev = lib.chub_set_path(path)
    try:
        got_u = _run("update", np.array(L, order=order), v, "torch")
        got_d = _run("downdate", np.array(L, order=order), vd, "torch")
    finally:
        lib.chub_set_path(prev)
    assert rel_fro(got_u, _oracle("update", L, v)[0]) <= tol_for(np.float64, n)
    assert rel_fro(got_d, _oracle("downdate", L, vd)[0]) <= tol_for(np.float64, n)


# ------------------------------------------------------------ reference property tests
@pytest.mark.parametrize("op", ["update", "downdate"])
@pytest.mark.parametrize("n", [10, 100, 700])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_upper_triangular_part_not_accessed(op, n, order, kind):
    # test_update.py:25-55 / test_downdate.py:28-58: junk (here NaN and random) above the diagonal
    L = synth_factor(n, 3)
    v = np.random.default_rng(4).standard_normal(n) if op == "update" else downdate_vector(L, 4)
    junk = np.triu(np.random.default_rng(5).standard_normal((n, n)), 1)
    junk[0, n - 1:] = np.nan if n > 1 else 0.0
    clean = _run(op, np.array(L, order=order), v, kind)
    dirty = _run(op, np.array(L + junk, order=order), v, kind)
    np.testing.assert_array_equal(np.triu(dirty, 1), junk)
    np.testing.assert_array_equal(np.tril(dirty), np.tril(clean))


@pytest.mark.parametrize("op", ["update", "downdate"])
@pytest.mark.parametrize("order", ["C", "F"])
def test_no_input_mutation(op, order):
    # test_update.py:58-85
    n = 100
    L = np.array(synth_factor(n, 8), order=order)
    v = np.random.default_rng(9).standard_normal(n) if op == "update" else downdate_vector(L, 9)
    fn = getattr(rank_1, op)
    for ow_L, ow_v in ((False, False), (True, False), (False, True)):
        Lc, vc = L.copy(order="K"), v.copy()
        out = fn(Lc, vc, overwrite_L=ow_L, overwrite_v=ow_v)
        if not ow_L:
            np.testing.assert_array_equal(Lc, L)
            assert out is not Lc
        else:
            assert out is Lc
        if not ow_v:
            np.testing.assert_array_equal(vc, v)
    Lt, vt = torch.from_numpy(L.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
    out = fn(Lt, vt)
    assert out is not Lt
    np.testing.assert_array_equal(Lt.cpu().numpy(), np.ascontiguousarray(L))
    np.testing.assert_array_equal(vt.cpu().numpy(), v)
    out = getattr(rank_1, op + "_in_place")(Lt, vt)
    assert out is Lt


@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_raise_on_zero_diagonal(kind):
    # test_update.py:129-141 / test_downdate.py:136-152; L must stay untouched
    for n in (5, 100, 700):
        L = synth_factor(n, 1)
        L[n // 2, n // 2] = 0.0
        v = np.random.default_rng(2).standard_normal(n)
        for op in ("update", "downdate"):
            if kind == "numpy":
                Lc, vc = L.copy(), v.copy()
                with pytest.raises(np.linalg.LinAlgError):
                    getattr(rank_1, op)(Lc, vc, overwrite_L=True, overwrite_v=True)
                np.testing.assert_array_equal(Lc, L)
            else:
                Lt, vt = torch.from_numpy(L.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
                with pytest.raises(np.linalg.LinAlgError):
                    getattr(rank_1, op)(Lt, vt, overwrite_L=True, overwrite_v=True)
                np.testing.assert_array_equal(Lt.cpu().numpy(), L)
                np.testing.assert_array_equal(vt.cpu().numpy(), v)


@pytest.mark.parametrize("kind", ["numpy", "torch"])
@pytest.mark.parametrize("n", [5, 100, 700, 1500])
@pytest.mark.parametrize("order", ["C", "F"])
def test_raise_on_indefinite_result(kind, n, order):
    # test_downdate.py:155-175: ||L^-1 v|| = 1.2 -> LinAlgError, L bit-unchanged, v <- p
    L = np.array(synth_factor(n, 11), order=order)
    v = downdate_vector(L, 12, norm_p=1.2)
    if kind == "numpy":
        Lc, vc = L.copy(order="K"), v.copy()
        with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
            rank_1.downdate(Lc, vc, overwrite_L=True, overwrite_v=True)
        np.testing.assert_array_equal(Lc, L)
        p = np.linalg.solve(np.tril(L), v)
        np.testing.assert_allclose(vc, p, rtol=1e-8, atol=1e-10)
    else:
        Lt, vt = torch.from_numpy(np.ascontiguousarray(L)).cuda(), torch.from_numpy(v.copy()).cuda()
        with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
            rank_1.downdate(Lt, vt, overwrite_L=True, overwrite_v=True)
        np.testing.assert_array_equal(Lt.cpu().numpy(), np.ascontiguousarray(L))
        p = np.linalg.solve(np.tril(L), v)  # v <- p = L^{-1} v (.pyx:239-259), on this path too
        np.testing.assert_allclose(vt.cpu().numpy(), p, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("n", [100, 1200])
def test_ill_conditioned_matrix(n):
    # test_update.py:144-166 / test_downdate.py:178-202
    g = np.random.default_rng(0)
    spectrum = np.sort(1.0 + g.gamma(10.0, 1.0, size=n))
    Q, _ = np.linalg.qr(g.standard_normal((n, n)))
    A = (Q * spectrum) @ Q.T
    A = 0.5 * (A + A.T)
    L = np.linalg.cholesky(A)
    v = Q[:, -1] * np.sqrt(spectrum[-1] * 99999.0)
    L_ud = rank_1.update(L, v)
    Aud = A + np.outer(v, v)
    assert np.linalg.norm(np.tril(L_ud) @ np.tril(L_ud).T - Aud) <= 1e-7 * np.linalg.norm(Aud)
    err = rel_fro(L_ud, _oracle("update", L, v)[0])
    assert err <= tol_for(np.float64, n), err
    v = Q[:, 0] * np.sqrt(spectrum[0] * (1.0 - 1e-6))
    L_dd = rank_1.downdate(L, v)
    Add = A - np.outer(v, v)
    assert np.linalg.norm(np.tril(L_dd) @ np.tril(L_dd).T - Add) <= 1e-7 * np.linalg.norm(Add)
    # near-singular downdate (rho^2 = 1e-6) against the oracle, at the same bar (the reference's own
    # Cython / Python / C-order variants differ by <= 1.5e-13 here)
    for order in ("C", "F"):
        for kind in ("numpy", "torch"):
            err = rel_fro(_run("downdate", np.array(L, order=order), v, kind), _oracle("downdate", L, v)[0])
            assert err <= tol_for(np.float64, n), (order, kind, err)


def test_negative_input_diagonal():
    # the sign fix (.pyx:110-113, :319-326): a factor with negative diagonal entries is still a
    # valid square root; the result has a positive diagonal and matches the oracle
    for n in (40, 700):
        L = synth_factor(n, 21)
        flip = np.random.default_rng(22).random(n) < 0.3
        L[:, flip] *= -1.0
        v = np.random.default_rng(23).standard_normal(n)
        got = rank_1.update(L, v)
        assert np.all(np.diag(got) > 0)
        assert rel_fro(got, _oracle("update", L, v)[0]) <= tol_for(np.float64, n)
        vd = downdate_vector(L, 24)
        got = rank_1.downdate(L, vd)
        assert np.all(np.diag(got) > 0)
        assert rel_fro(got, _oracle("downdate", L, vd)[0]) <= tol_for(np.float64, n)


# ---------------------------------------------------------------------- batched
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("layout", ["C", "F"])
def test_batched_matches_oracle_loop(dtype, layout):
    B, n = 24, 256
    Ls = np.stack([synth_factor(n, [0, b], dtype) for b in range(B)])
    vs = np.stack([np.random.default_rng([1, b]).standard_normal(n).astype(dtype) for b in range(B)])
    Lt = torch.from_numpy(Ls).cuda()
    if layout == "F":
        Lt = Lt.transpose(1, 2).contiguous().transpose(1, 2)
    out = rank_1.update_batched(Lt, torch.from_numpy(vs).cuda())
    got = out.cpu().numpy()
    for b in range(B):
        assert rel_fro(got[b], _oracle("update", Ls[b], vs[b])[0]) <= tol_for(dtype, n)
    vd = np.stack([downdate_vector(Ls[b], [2, b]) for b in range(B)])
    vd[3] = downdate_vector(Ls[3], 5, norm_p=1.2)  # one factor loses positive definiteness
    Ls[7, 9, 9] = 0.0  # and one has a zero on its diagonal
    Lt = torch.from_numpy(Ls).cuda()
    out, status = rank_1.downdate_batched(Lt, torch.from_numpy(vd).cuda(), raise_on_error=False)
    status = status.cpu().numpy()
    assert status[3] == 2 and status[7] == 1 and np.count_nonzero(status) == 2
    got = out.cpu().numpy()
    np.testing.assert_array_equal(got[3], Ls[3])
    np.testing.assert_array_equal(got[7], Ls[7])
    slack = 1 if dtype == np.float64 else 30
    for b in (0, 1, 23):
        assert rel_fro(got[b], _oracle("downdate", Ls[b], vd[b])[0]) <= tol_for(dtype, n) * slack
    with pytest.raises(np.linalg.LinAlgError):
        rank_1.downdate_batched(Lt, torch.from_numpy(vd).cuda())


# ------------------------------------------------------------------- larger sizes
@pytest.mark.parametrize("order", ["C", "F"])
def test_update_n4096_matches_oracle(order):
    n = 4096
    L = synth_factor(n, 0)
    v = np.random.default_rng(1).standard_normal(n)
    L_ref, v_ref = _oracle("update", L, v)
    got = _run("update", np.array(L, order=order), v, "torch")
    assert rel_fro(got, L_ref) <= tol_for(np.float64, n)
    got = _run("update", np.array(L, order=order), v, "numpy")
    assert rel_fro(got, L_ref) <= tol_for(np.float64, n)


@pytest.mark.parametrize("order", ["C", "F"])
def test_downdate_n4096_matches_oracle(order):
    n = 4096
    L = synth_factor(n, 0)
    v = downdate_vector(L, 2, norm_p=np.sqrt(1 - 1e-6))  # valid but near-singular (config 3 (i))
    L_ref, _ = _oracle("downdate", L, v)
    got = _run("downdate", np.array(L, order=order), v, "torch")
    assert rel_fro(got, L_ref) <= tol_for(np.float64, n)


# ------------------------------------------------ block-row-cyclic driver on one GPU (world = 1)
@pytest.mark.parametrize("n", [700, 1000, 2048])
def test_rowcyclic_single_rank_matches_oracle(n):
    from cholupdates_b200.distributed import RowCyclicFactor

    L = synth_factor(n, 31)
    L[:, 3] *= -1.0
    A = torch.from_numpy(L.copy()).cuda()
    fac = RowCyclicFactor(A, n)
    L_ref = np.array(L, order="F")
    for u in range(2):
        v = np.random.default_rng([7, u]).standard_normal(n)
        fac.update(torch.from_numpy(v.copy()).cuda(), per_panel=(u == 0))  # both single-update forms
        oracle.update(L_ref, v.copy())
    got = A.cpu().numpy()
    assert rel_fro(got, L_ref) <= tol_for(np.float64, n)
    np.testing.assert_array_equal(np.triu(got, 1), 0)
    Lz = torch.from_numpy(L.copy()).cuda()
    Lz[n // 2, n // 2] = 0.0
    keep = Lz.clone()
    with pytest.raises(np.linalg.LinAlgError):
        RowCyclicFactor(Lz, n).update(torch.ones(n, dtype=torch.float64, device="cuda"))
    assert torch.equal(keep, Lz)


@pytest.mark.parametrize("fuse", [None, 1, 2])
@pytest.mark.parametrize("n,K", [(300, 1), (700, 5), (1000, 2), (2048, 11)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rowcyclic_wavefront_single_rank_matches_oracle(n, K, dtype, fuse):
    """K successive updates as one wavefront (anti-diagonals of (update, panel) tasks) == K oracle calls;
    the scratch vectors are rotg's z_k, as the Cython path leaves them in v."""
    from cholupdates_b200.distributed import RowCyclicFactor

    L = synth_factor(n, 33).astype(dtype)
    L[:, 3] *= -1.0
    V = np.stack([np.random.default_rng([9, u]).standard_normal(n) for u in range(K)]).astype(dtype)
    A = torch.from_numpy(L.copy()).cuda()
    A_upper = torch.triu(torch.full_like(A, float("nan")), 1)
    A += A_upper  # the strict upper triangle must never reach the arithmetic
    # fuse = 2: consecutive updates share one pass over every panel (odd K: the last group is short)
    Z = RowCyclicFactor(A, n).update_many(torch.from_numpy(V.copy()).cuda(), fuse=fuse)
    L_ref = np.array(L, order="F")
    z_ref = []
    for u in range(K):
        vu = V[u].copy()
        oracle.update(L_ref, vu)
        z_ref.append(vu)
    got = A.cpu().numpy()
    assert rel_fro(got, L_ref) <= tol_for(dtype, n) * (1 if dtype == np.float64 else K)
    assert np.all(np.isnan(got[np.triu_indices(n, 1)]))
    if dtype == np.float64:
        np.testing.assert_allclose(Z.cpu().numpy(), np.stack(z_ref), rtol=1e-8, atol=1e-10)
    Lz = torch.from_numpy(L.copy()).cuda()
    Lz[n // 2, n // 2] = 0.0
    keep = Lz.clone()
    with pytest.raises(np.linalg.LinAlgError):
        RowCyclicFactor(Lz, n).update_many(torch.ones((2, n), dtype=Lz.dtype, device="cuda"))
    assert torch.equal(keep, Lz)


@pytest.mark.parametrize("per_panel", [None, True])
@pytest.mark.parametrize("n", [300, 1000, 2048])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rowcyclic_downdate_single_rank_matches_oracle(n, dtype, per_panel):
    """Distributed downdate driver on one rank: panel-wise forward substitution (on the wavefront kernels,
    or per panel), rho^2 test, local sweep."""
    from cholupdates_b200.distributed import RowCyclicFactor

    L = synth_factor(n, 35).astype(dtype)
    L[:, 3] *= -1.0  # a negative input diagonal entry: the column's sign flip must survive (.pyx:319-326)
    v = downdate_vector(L, 36)
    A = torch.from_numpy(L.copy()).cuda()
    A += torch.triu(torch.full_like(A, float("nan")), 1)
    z = RowCyclicFactor(A, n).downdate(torch.from_numpy(v.copy()).cuda(), per_panel=per_panel)
    L_ref, z_ref = np.array(L, order="F"), v.copy()
    oracle.downdate(L_ref, z_ref)
    got = A.cpu().numpy()
    assert rel_fro(got, L_ref) <= tol_for(dtype, n) * (1 if dtype == np.float64 else 30)
    assert np.all(np.isnan(got[np.triu_indices(n, 1)]))
    if dtype == np.float64:
        np.testing.assert_allclose(z.cpu().numpy(), z_ref, rtol=1e-8, atol=1e-10)
    # ||L^{-1} v|| = 1.2: LinAlgError, factor untouched (test_downdate.py:155-175)
    keep = A.clone()
    bad = downdate_vector(np.tril(L_ref).astype(dtype), 37, norm_p=1.2)
    with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
        RowCyclicFactor(A, n).downdate(torch.from_numpy(bad).cuda(), per_panel=per_panel)
    assert torch.equal(torch.tril(keep), torch.tril(A))


def test_sharded_batched_single_process():
    from cholupdates_b200.distributed import shard_bounds, sharded_batched

    B, n = 6, 64
    Ls = np.stack([synth_factor(n, [0, b]) for b in range(B)])
    vs = np.stack([np.random.default_rng([1, b]).standard_normal(n) for b in range(B)])
    lo, hi = shard_bounds(B, 0, 1)
    out, status = sharded_batched("update", torch.from_numpy(Ls[lo:hi]).cuda(), torch.from_numpy(vs[lo:hi]).cuda())
    assert status.tolist() == [0] * B
    for b in range(B):
        assert rel_fro(out[b].cpu().numpy(), _oracle("update", Ls[b], vs[b])[0]) <= tol_for(np.float64, n)


@pytest.mark.parametrize("op", ["update", "downdate"])
@pytest.mark.parametrize("n", [100, 600, 1500])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_nan_filled_upper_triangle(op, n, order, dtype):
    """Every entry above the diagonal is NaN: the result below must be bit-identical to the clean run
    and the NaNs must come back untouched (stronger form of test_update.py:25-55; regression test for
    a 0 * NaN leak in the pipelined downdate sweep)."""
    L = synth_factor(n, 17, dtype)
    v = (np.random.default_rng(18).standard_normal(n).astype(dtype) if op == "update"
         else downdate_vector(L, 18))
    dirty_in = L.copy()
    dirty_in[np.triu_indices(n, 1)] = np.nan
    for kind in ("torch", "numpy"):
        clean = _run(op, np.array(L, order=order), v, kind)
        dirty = _run(op, np.array(dirty_in, order=order), v, kind)
        assert np.all(np.isnan(dirty[np.triu_indices(n, 1)]))
        np.testing.assert_array_equal(np.tril(dirty), np.tril(clean))
        assert not np.any(np.isnan(np.tril(dirty)))


@pytest.mark.parametrize("n", [600, 1536])
def test_native_column_major_kernel(n, monkeypatch):
    """Column-major factors normally go through a row-major scratch copy; the native column-major
    dataflow kernel (df_kernel_v2) is the fallback when that scratch cannot be allocated.  Force it."""
    monkeypatch.setenv("CHUB_COL_MAJOR_NATIVE", "1")
    L = synth_factor(n, 41)
    v = np.random.default_rng(42).standard_normal(n)
    vd = downdate_vector(L, 43)
    got_u = _run("update", np.asfortranarray(L), v, "torch")
    got_d = _run("downdate", np.asfortranarray(L), vd, "torch")
    assert rel_fro(got_u, _oracle("update", L, v)[0]) <= tol_for(np.float64, n)
    assert rel_fro(got_d, _oracle("downdate", L, vd)[0]) <= tol_for(np.float64, n)


@pytest.mark.parametrize("n,ld", [(100, 128), (700, 1024), (1000, 1002)])
@pytest.mark.parametrize("op", ["update", "downdate"])
def test_c_abi_leading_dimension_larger_than_n(n, ld, op):
    """The C ABI takes a leading dimension: an n x n factor living in the corner of a wider row-major
    buffer (a superset of the reference, which only accepts contiguous arrays, _seeger.py:260-264).
    Everything outside the n x n lower triangle must come back bit-identical."""
    import ctypes

    lib = _lib.load()
    L = synth_factor(n, 51)
    v = np.random.default_rng(52).standard_normal(n) if op == "update" else downdate_vector(L, 52)
    big = np.random.default_rng(53).standard_normal((n + 3, ld))
    big[:n, :n] = np.tril(L) + np.triu(big[:n, :n], 1)
    bt = torch.from_numpy(big.copy()).cuda()
    vt = torch.from_numpy(v.copy()).cuda()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = getattr(lib, f"chub_{op}_f64")(bt.data_ptr(), n, ld, 0, vt.data_ptr(), 1, None, 0,
                                        status.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0 and int(status.item()) == 0
    got = bt.cpu().numpy()
    assert rel_fro(got[:n, :n], _oracle(op, L, v)[0]) <= tol_for(np.float64, n)
    mask = np.ones_like(big, dtype=bool)
    mask[:n, :n] = np.triu(np.ones((n, n), dtype=bool), 1)
    np.testing.assert_array_equal(got[mask], big[mask])


# ------------------------------------------------------------------- config 1 (the notebook workload)
@pytest.mark.parametrize("n", [1000, 5000])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("inplace", [False, True])
def test_config1_notebook_inputs(n, order, inplace):
    """BASELINE.json configs[0] / SURVEY 8d: random_spd_matrix(fast=True) + scipy cholesky + v ~ N(0, 10^2),
    check_diag=False, NumPy in / NumPy out, copy and in-place (notebooks/benchmark_rank_1.ipynb:119-180)."""
    L0, v0 = config1_inputs(n)
    L_ref, v_ref = _oracle("update", L0, v0)
    L_in, v_in = np.array(L0, order=order), v0.copy()
    out = rank_1.update(L_in, v_in, check_diag=False, overwrite_L=inplace, overwrite_v=inplace)
    assert (out is L_in) == inplace
    assert out.flags.c_contiguous == L_in.flags.c_contiguous and out.flags.f_contiguous == L_in.flags.f_contiguous
    assert rel_fro(out, L_ref) <= tol_for(np.float64, n), rel_fro(out, L_ref)
    np.testing.assert_array_equal(np.triu(out, 1), 0)
    if inplace:
        assert rel_vec(v_in, v_ref) <= tol_for(np.float64, n), rel_vec(v_in, v_ref)
    else:
        np.testing.assert_array_equal(L_in, np.array(L0, order=order))
        np.testing.assert_array_equal(v_in, v0)
    # and as a CUDA tensor
    got = _run("update", np.array(L0, order=order), v0, "torch", check_diag=False)
    assert rel_fro(got, L_ref) <= tol_for(np.float64, n)


# ------------------------------------------------------------------- the batched config at full size
@pytest.mark.parametrize("op", ["update", "downdate"])
def test_batched_8192x256_sampled(op):
    """BASELINE.json configs[3] at its full size: 8192 independent n = 256 factors; every status is 0 and a
    sample of the factors (first / middle / last CTAs of the grid) matches the oracle."""
    B, n = 8192, 256
    g = torch.Generator(device="cuda")
    g.manual_seed(77)
    L = torch.tril(torch.randn((B, n, n), dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n), -1)
    L.diagonal(dim1=1, dim2=2).copy_(1.0 + torch.rand((B, n), dtype=torch.float64, device="cuda", generator=g))
    if op == "update":
        v = torch.randn((B, n), dtype=torch.float64, device="cuda", generator=g)
    else:
        u = torch.randn((B, n), dtype=torch.float64, device="cuda", generator=g)
        u = 0.8 * u / u.norm(dim=1, keepdim=True)
        v = torch.bmm(L, u.unsqueeze(2)).squeeze(2)  # p = 0.8 u
    sample = sorted(set(list(range(6)) + list(range(4093, 4099)) + list(range(B - 6, B)) + [1000, 2047, 2048, 6000]))
    L0, v0 = L[sample].cpu().numpy(), v[sample].cpu().numpy()
    out, status = getattr(rank_1, op + "_batched")(L, v, check_diag=True, overwrite_L=True, overwrite_v=True,
                                                   raise_on_error=False)
    assert out is L
    assert int(torch.count_nonzero(status).item()) == 0
    got = L[sample].cpu().numpy()
    for i in range(len(sample)):
        L_ref, _ = _oracle(op, L0[i], v0[i])
        assert rel_fro(got[i], L_ref) <= tol_for(np.float64, n), (sample[i],)
    assert bool(torch.isfinite(torch.tril(L)).all())


# ------------------------------------------------------------------- resident-factor handle
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("n", [100, 1500, 3000])
def test_resident_factor_successive_updates(order, n):
    from cholupdates_b200.resident import ResidentFactor

    L = synth_factor(n, 61)
    junk = np.triu(np.random.default_rng(62).standard_normal((n, n)), 1)
    host = np.array(L + junk, order=order)
    L_ref = np.array(L, order="F")
    with ResidentFactor(host) as fac:
        for u in range(3):
            v = np.random.default_rng([63, u]).standard_normal(n)
            z_ref = v.copy()
            oracle.update(L_ref, z_ref)
            z = fac.update(v)
            assert rel_vec(z, z_ref) <= tol_for(np.float64, n)
        vd = downdate_vector(np.tril(L_ref), 64)
        oracle.downdate(L_ref, vd.copy())
        fac.downdate(vd)
        bad = downdate_vector(np.tril(L_ref), 65, norm_p=1.2)
        with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
            fac.downdate(bad)
        out = fac.download(out=host)
    assert out is host
    assert rel_fro(host, L_ref) <= tol_for(np.float64, n) * 2
    np.testing.assert_array_equal(np.triu(host, 1), junk)  # the upper triangle never travelled
    fresh = None
    with ResidentFactor(np.array(L, order=order)) as fac:
        fresh = fac.download()
    np.testing.assert_array_equal(fresh, np.array(np.tril(L), order=order))
    assert fresh.flags.f_contiguous == (order == "F") or n == 1


# ------------------------------------------------------------------- strided views (ld > n) through the front-end
@pytest.mark.parametrize("op", ["update", "downdate"])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("kind", ["numpy", "torch"])
@pytest.mark.parametrize("n,ld", [(100, 128), (1000, 1024), (700, 702)])
def test_strided_view_through_front_end(op, order, kind, n, ld):
    """`strided=True`: an n x n factor living in the corner of a wider buffer (ld > n), both orders, NumPy
    and CUDA tensors; everything outside the n x n lower triangle comes back bit-identical.  Without the
    flag the reference's ValueError is kept (_seeger.py:260-264)."""
    L = synth_factor(n, 71)
    v = np.random.default_rng(72).standard_normal(n) if op == "update" else downdate_vector(L, 72)
    # a wider buffer whose memory order is `order` and whose leading dimension is `ld`
    shape = (n + 3, ld) if order == "C" else (ld, n + 3)
    big = np.array(np.random.default_rng(73).standard_normal(shape), order=order)
    big[:n, :n] = np.tril(L) + np.triu(big[:n, :n], 1)
    fn = getattr(rank_1, op)
    L_ref, _ = _oracle(op, L, v)
    if kind == "numpy":
        buf = big.copy(order="K")
        view = buf[:n, :n]
        mk_v = lambda: v.copy()  # noqa: E731
    else:
        store = torch.from_numpy(np.ascontiguousarray(big if order == "C" else big.T)).cuda()
        buf = store if order == "C" else store.t()
        view = buf[:n, :n]
        mk_v = lambda: torch.from_numpy(v.copy()).cuda()  # noqa: E731
    with pytest.raises(ValueError):
        fn(view, mk_v(), overwrite_L=True)
    out = fn(view, mk_v(), overwrite_L=True, overwrite_v=True, strided=True)
    assert out is view
    got = buf if kind == "numpy" else buf.cpu().numpy()
    assert rel_fro(got[:n, :n], L_ref) <= tol_for(np.float64, n)
    mask = np.ones(big.shape, dtype=bool)
    mask[:n, :n] = np.triu(np.ones((n, n), dtype=bool), 1)
    np.testing.assert_array_equal(got[mask], big[mask])
    # copy mode keeps the input and returns a dense result
    before = np.array(got, copy=True)
    v2 = 0.3 * v  # (a second downdate with the full v could leave the positive-definite cone)
    res = fn(view, v2.copy() if kind == "numpy" else torch.from_numpy(v2.copy()).cuda(), strided=True)
    res = res if kind == "numpy" else res.cpu().numpy()
    np.testing.assert_array_equal(buf if kind == "numpy" else buf.cpu().numpy(), before)
    assert rel_fro(res, _oracle(op, np.tril(before[:n, :n]), v2)[0]) <= tol_for(np.float64, n) * 2


def test_dlpack_input_is_viewed_zero_copy():
    """An array type that is neither numpy nor torch but speaks DLPack (CuPy ...) is updated in place
    through a zero-copy torch view; a stand-in class exercises the protocol without CuPy."""

    class Foreign:
        def __init__(self, t):
            self.t = t
            self.ndim, self.shape, self.dtype = t.ndim, tuple(t.shape), t.dtype

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    n = 300
    L = synth_factor(n, 81)
    v = np.random.default_rng(82).standard_normal(n)
    Lt, vt = torch.from_numpy(L.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
    fl, fv = Foreign(Lt), Foreign(vt)
    out = rank_1.update(fl, fv, overwrite_L=True, overwrite_v=True)
    assert out is fl
    assert rel_fro(Lt.cpu().numpy(), _oracle("update", L, v)[0]) <= tol_for(np.float64, n)


# ------------------------------------------------------------------- method="cho_factor" and the device generators
@pytest.mark.parametrize("n", [5, 100, 600])
@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_cho_factor_method_cross_check(n, kind):
    """The reference's own cross-check (update/conftest.py:27-41 enrols every method): the O(N^3)
    re-factorisation and the Seeger path give the same factor; the strict upper triangle is restored."""
    A, L = spd_factor(n, 91)
    junk = np.triu(np.random.default_rng(92).standard_normal((n, n)), 1)
    v = np.random.default_rng(93).normal(scale=10, size=n)
    vd = downdate_vector(L, 94)
    for op, vec in (("update", v), ("downdate", vd)):
        a = _run(op, L + junk, vec, kind, method="cho_factor")
        b = _run(op, L + junk, vec, kind, method="seeger")
        assert rel_fro(a, b) <= 1e-10
        np.testing.assert_array_equal(np.triu(a, 1), junk)
    with pytest.raises(np.linalg.LinAlgError):
        rank_1.downdate(L, downdate_vector(L, 95, norm_p=1.2), method="cho_factor")


def test_device_generators():
    from cholupdates_b200 import utils

    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    for fast in (True, False):
        M = utils.random_spd_matrix(200, fast=fast, generator=g)
        assert torch.equal(M, M.T)
        assert float(torch.linalg.eigvalsh(M).min()) > 0  # (tests/cholupdates/test_utils.py:24-33)
    spec, Q = utils.random_spd_eigendecomposition(50, generator=g)
    assert float(spec.min()) >= 1.0 and bool((spec[1:] >= spec[:-1]).all())
    assert float((Q.T @ Q - torch.eye(50, dtype=Q.dtype, device=Q.device)).abs().max()) < 1e-12
    L = torch.linalg.cholesky(utils.random_spd_matrix(300, generator=g))
    v = utils.random_rank_1_downdate(L, generator=g)
    p = torch.linalg.solve_triangular(L, v[:, None], upper=False)[:, 0]
    assert 0.2 <= float(p @ p) <= 0.9
    L_dd = rank_1.downdate(L, v)
    assert bool((torch.diagonal(L_dd) > 0).all())


# ------------------------------------------------- the A/B switches of the persistent kernels stay correct
@pytest.mark.parametrize(
    "env",
    [
        {"CHUB_RM5": "0"},  # first-generation row warps
        {"CHUB_RM5_3D": "0"},  # second generation without the 3-D tensor map (2-D boxes only)
        {"CHUB_DD_SWEEP_V2": "1", "CHUB_DD_COEF_V1": "1"},  # first-generation downdate sweep / coefficient kernel
        {"CHUB_DD_SWEEP_V1": "1"},  # launch-per-row-block sweep
    ],
)
@pytest.mark.parametrize("n", [1022, 3000])
def test_persistent_kernel_switches(n, env, monkeypatch):
    """The library reads its A/B switches at call time: every variant of the row-major persistent path (and its
    ragged right edge, n = 1022: the last 16-column block is outside the 3-D view) against the oracle."""
    for k, val in env.items():
        monkeypatch.setenv(k, val)
    L = synth_factor(n, n + 7, np.float64)
    v = np.random.default_rng(n).standard_normal(n)
    L_ref, v_ref = _oracle("update", L, v)
    Lt = torch.from_numpy(L.copy()).cuda()
    Lt = Lt + torch.triu(torch.full_like(Lt, float("nan")), 1)
    vt = torch.from_numpy(v.copy()).cuda()
    rank_1.update(Lt, vt, check_diag=False, overwrite_L=True, overwrite_v=True)
    got = Lt.cpu().numpy()
    assert np.all(np.isnan(got[np.triu_indices(n, 1)]))
    assert rel_fro(np.nan_to_num(got), L_ref) <= tol_for(np.float64, n)
    assert rel_vec(vt.cpu().numpy(), v_ref) <= tol_for(np.float64, n)
    vd = downdate_vector(np.tril(L_ref), n + 3)
    L_dd, vd_ref = _oracle("downdate", np.tril(L_ref), vd)
    Ld = torch.from_numpy(np.tril(L_ref).copy()).cuda()
    vdt = torch.from_numpy(vd.copy()).cuda()
    rank_1.downdate(Ld, vdt, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Ld.cpu().numpy(), L_dd) <= tol_for(np.float64, n)
    assert rel_vec(vdt.cpu().numpy(), vd_ref) <= tol_for(np.float64, n)


# ------------------------------------------------- every host path (NumPy in / NumPy out) stays correct
@pytest.mark.parametrize(
    "env",
    [
        {},  # staged (default for pageable arrays)
        {"CHUB_HOST_PATH": "streamed"},  # streamed panels + pinned bounce ring (updates)
        {"CHUB_HOST_PATH": "direct"},  # plain 2-D copies
        {"CHUB_HOST_STAGE_MAX_MB": "1"},  # staging buffer too small: streamed (update) / direct (downdate)
        {"CHUB_HOST_BLOCKS": "1"},  # staged, a single block
        {"CHUB_HOST_BLOCKS": "97"},  # staged, many ragged blocks
    ],
)
@pytest.mark.parametrize("order", ["C", "F"])
def test_host_paths(env, order, monkeypatch):
    """Pageable NumPy arrays through each host path: lower triangle vs the oracle, junk above the diagonal
    untouched, copy mode returns a fresh array of the same order, the not-positive-definite downdate leaves L
    as it was and v <- p."""
    for k, val in env.items():
        monkeypatch.setenv(k, val)
    n = 3000
    base = synth_factor(n, 91)
    junk = np.triu(np.random.default_rng(92).standard_normal((n, n)), 1)
    L = np.array(base + junk, order=order)
    v = np.random.default_rng(93).standard_normal(n)
    L_ref, v_ref = _oracle("update", base, v)
    out = rank_1.update(L, v.copy(), check_diag=False)  # copy mode
    assert out is not L and out.flags.f_contiguous == L.flags.f_contiguous
    assert np.array_equal(np.triu(out, 1), junk) and np.array_equal(np.triu(L, 1), junk)
    assert rel_fro(np.tril(out), L_ref) <= tol_for(np.float64, n)
    vv = v.copy()
    rank_1.update(L, vv, check_diag=True, overwrite_L=True, overwrite_v=True)  # in place
    assert np.array_equal(np.triu(L, 1), junk)
    assert rel_fro(np.tril(L), L_ref) <= tol_for(np.float64, n)
    assert rel_vec(vv, v_ref) <= tol_for(np.float64, n)
    vd = downdate_vector(np.tril(L_ref), 94)
    L_dd, vd_ref = _oracle("downdate", np.tril(L_ref), vd)
    vdd = vd.copy()
    rank_1.downdate(L, vdd, overwrite_L=True, overwrite_v=True)
    assert np.array_equal(np.triu(L, 1), junk)
    assert rel_fro(np.tril(L), L_dd) <= tol_for(np.float64, n) * 2
    assert rel_vec(vdd, vd_ref) <= tol_for(np.float64, n) * 2
    # ||L^{-1} v|| = 1.2: LinAlgError, L bit-unchanged, v <- p
    before = L.copy(order="K")
    p = np.random.default_rng(95).standard_normal(n)
    p *= 1.2 / np.linalg.norm(p)
    vbad = np.tril(L) @ p
    with pytest.raises(np.linalg.LinAlgError):
        rank_1.downdate(L, vbad, overwrite_L=True, overwrite_v=True)
    assert np.array_equal(L, before)
    assert rel_vec(vbad, p) <= 1e-9
