"""Full-size checks at BASELINE.json's configuration 2 / 3 (n = 16384, fp64) on the GPU.

The oracle (compiled C, column-major fast path) finishes n = 16384 in a fraction of a second, so
the comparison is direct; in addition a size-independent property is checked:
(L' L'^T) x = L (L^T x) +/- v (v^T x) for random x, which needs only mat-vecs.
"""

import numpy as np
import pytest

import seeger_oracle as oracle
from helpers import downdate_vector, rel_fro, synth_factor, tol_for

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import cholupdates_b200  # noqa: E402

rank_1 = cholupdates_b200.rank_1
N = 16384


@pytest.fixture(scope="module")
def big():
    L = synth_factor(N, 0)
    v = np.random.default_rng(1).standard_normal(N)
    return L, v


def _to_cuda(L, order):
    if order == "C":
        return torch.from_numpy(L).cuda()
    return torch.from_numpy(np.ascontiguousarray(L.T)).cuda().t()


@pytest.mark.parametrize("order", ["C", "F"])
def test_update_16384(big, order):
    L, v = big
    L_ref, v_ref = np.array(L, order="F"), v.copy()
    oracle.update(L_ref, v_ref)
    Lt = _to_cuda(L, order)
    Lt = Lt + torch.triu(torch.full_like(Lt, float("nan")), 1)  # NaN above the diagonal
    vt = torch.from_numpy(v.copy()).cuda()
    out = rank_1.update(Lt, vt, check_diag=False, overwrite_L=True, overwrite_v=True)
    assert out is Lt
    got = Lt.cpu().numpy()
    assert np.all(np.isnan(got[np.triu_indices(N, 1)]))  # never touched
    assert rel_fro(np.nan_to_num(got), L_ref) <= tol_for(np.float64, N)
    err_v = np.linalg.norm(vt.cpu().numpy() - v_ref) / np.linalg.norm(v_ref)
    assert err_v <= tol_for(np.float64, N), err_v  # v: the same bar as L (north star)
    # size-independent property
    Lo = torch.tril(torch.nan_to_num(Lt))
    x = torch.from_numpy(np.random.default_rng(3).standard_normal(N)).cuda()
    Lold = torch.from_numpy(np.tril(L)).cuda()
    vv = torch.from_numpy(v).cuda()
    lhs = Lo @ (Lo.T @ x)
    rhs = Lold @ (Lold.T @ x) + vv * (vv @ x)
    assert float(torch.linalg.norm(lhs - rhs) / torch.linalg.norm(rhs)) <= 1e-11


def test_downdate_16384_near_singular_and_error_paths(big):
    L, _ = big
    # (i) valid but near-singular: alpha^2 = 1 - 1e-6 (mirrors test_downdate.py:189-190)
    v = downdate_vector(L, 2, norm_p=np.sqrt(1 - 1e-6))
    L_ref, v_ref = np.array(L, order="F"), v.copy()
    oracle.downdate(L_ref, v_ref)
    Lt, vt = torch.from_numpy(L.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
    rank_1.downdate(Lt, vt, check_diag=True, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Lt.cpu().numpy(), L_ref) <= tol_for(np.float64, N)
    # (ii) invalid: alpha = 1.2 and alpha^2 = 1 + 1e-6 -> LinAlgError, L bit-unchanged
    for norm_p in (1.2, np.sqrt(1 + 1e-6)):
        v = downdate_vector(L, 2, norm_p=norm_p)
        Lt, vt = torch.from_numpy(L.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
        with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
            rank_1.downdate(Lt, vt, overwrite_L=True, overwrite_v=True)
        assert torch.equal(Lt.cpu(), torch.from_numpy(L))
    # (iii) planted zero on the diagonal with check_diag=True
    Lz = L.copy()
    Lz[12345, 12345] = 0.0
    Lt, vt = torch.from_numpy(Lz.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
    with pytest.raises(np.linalg.LinAlgError, match="zeros on its diagonal"):
        rank_1.downdate(Lt, vt, overwrite_L=True, overwrite_v=True)
    assert torch.equal(Lt.cpu(), torch.from_numpy(Lz))


@pytest.mark.parametrize("n", [20000])
@pytest.mark.parametrize("order", ["C", "F"])
def test_update_and_downdate_above_16464(n, order):
    """Sizes beyond what one generation of row warps covers (147 workers x 14 warps x 8 rows = 16464):
    the persistent kernel deals further row groups to the same warps."""
    L = synth_factor(n, 5)
    v = np.random.default_rng(6).standard_normal(n)
    L_ref, v_ref = np.array(L, order="F"), v.copy()
    oracle.update(L_ref, v_ref)
    Lt = _to_cuda(L, order)
    Lt = Lt + torch.triu(torch.full_like(Lt, float("nan")), 1)
    vt = torch.from_numpy(v.copy()).cuda()
    rank_1.update(Lt, vt, check_diag=False, overwrite_L=True, overwrite_v=True)
    got = Lt.cpu().numpy()
    assert np.all(np.isnan(got[np.triu_indices(n, 1)]))
    assert rel_fro(np.nan_to_num(got), L_ref) <= tol_for(np.float64, n)
    assert np.linalg.norm(vt.cpu().numpy() - v_ref) <= tol_for(np.float64, n) * np.linalg.norm(v_ref)
    del got
    vd = downdate_vector(np.tril(L_ref), 7)
    L_dd = np.array(L_ref, order="F")
    oracle.downdate(L_dd, vd.copy())
    Lt = torch.nan_to_num(Lt)
    rank_1.downdate(Lt, torch.from_numpy(vd.copy()).cuda(), check_diag=True, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Lt.cpu().numpy(), L_dd) <= tol_for(np.float64, n) * 2


def test_blocked_driver_column_major_vs_oracle(monkeypatch):
    """Column-major factors beyond the native kernel's row capacity (n > 28224) go through diagonal blocks +
    rect_apply_cm_kernel.  The knob sends n = 20000 down that path (two diagonal blocks of ~10000), where the
    oracle is still a direct comparison."""
    monkeypatch.setenv("CHUB_CM_NATIVE_MAX_N", "16384")
    n = 20000
    L = synth_factor(n, 15)
    v = np.random.default_rng(16).standard_normal(n)
    L_ref, v_ref = np.array(L, order="F"), v.copy()
    oracle.update(L_ref, v_ref)
    Lt = _to_cuda(L, "F")
    vt = torch.from_numpy(v.copy()).cuda()
    rank_1.update(Lt, vt, check_diag=False, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Lt.cpu().numpy(), L_ref) <= tol_for(np.float64, n)
    assert np.linalg.norm(vt.cpu().numpy() - v_ref) <= tol_for(np.float64, n) * np.linalg.norm(v_ref)
    vd = downdate_vector(np.tril(L_ref), 17)
    L_dd, vd_ref = np.array(L_ref, order="F"), vd.copy()
    oracle.downdate(L_dd, vd_ref)
    vdt = torch.from_numpy(vd.copy()).cuda()
    rank_1.downdate(Lt, vdt, check_diag=True, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Lt.cpu().numpy(), L_dd) <= tol_for(np.float64, n) * 2
    assert np.linalg.norm(vdt.cpu().numpy() - vd_ref) <= tol_for(np.float64, n) * 2 * np.linalg.norm(vd_ref)


@pytest.mark.parametrize("order", ["C", "F"])
def test_update_and_downdate_n32768_properties(order):
    """n = 32768 (8.6 GB, two diagonal blocks + one rectangle in either layout), checked without a host copy:
    the leading 4096 x 4096 block against the oracle (the update of a leading block depends on that block and
    v[:m] only), and 64 sampled rows of L' L'^T against L L^T +/- v v^T (all rows enter through the products)."""
    n, m = 32768, 4096
    g = torch.Generator(device="cuda").manual_seed(32768)
    L = torch.tril(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n), -1)
    L.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
    if order == "F":
        L = L.t().contiguous().t()
    v = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    rows = torch.randint(0, n, (64,), device="cuda", generator=g)
    rows[0], rows[1] = n - 1, n // 2 + 5
    A_rows = L[rows] @ L.t()
    lead, v_lead = np.array(L[:m, :m].cpu().numpy(), order="F"), v[:m].cpu().numpy().copy()
    for op, sign in (("update", 1.0), ("downdate", -1.0)):
        if op == "downdate":  # v = L p with ||p|| = 0.7 keeps the result positive definite
            p = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
            v = L @ (0.7 * p / p.norm())
            v_lead = v[:m].cpu().numpy().copy()
        want = A_rows + sign * torch.outer(v[rows], v)
        v_in = v.clone()
        getattr(rank_1, op)(L, v_in, check_diag=False, overwrite_L=True, overwrite_v=True)
        assert bool((torch.triu(L, 1) == 0).all())
        A_rows = L[rows] @ L.t()
        err = float((A_rows - want).norm() / want.norm())
        assert err <= 1e-12 * np.sqrt(n), (op, err)
        getattr(oracle, op)(lead, v_lead)
        assert rel_fro(L[:m, :m].cpu().numpy(), lead) <= tol_for(np.float64, n), op
        assert np.linalg.norm(v_in[:m].cpu().numpy() - v_lead) <= tol_for(np.float64, n) * np.linalg.norm(v_lead), op


@pytest.mark.parametrize("blocked", [False, True])
def test_column_major_fp32_n20000(blocked, monkeypatch):
    """fp32, column-major, n = 20000: the two-warps-per-block-row workers (Q = 5) and, with the knob, the blocked
    driver (two diagonal blocks + rect_apply_cm_kernel<float>), update and downdate against the fp32 oracle."""
    if blocked:
        monkeypatch.setenv("CHUB_CM_NATIVE_MAX_N", "16384")
    n = 20000
    L = synth_factor(n, 25, np.float32)
    v = np.random.default_rng(26).standard_normal(n).astype(np.float32)
    L_ref, v_ref = np.array(L, order="F"), v.copy()
    oracle.update(L_ref, v_ref)
    Lt = _to_cuda(L, "F")
    vt = torch.from_numpy(v.copy()).cuda()
    rank_1.update(Lt, vt, check_diag=False, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Lt.cpu().numpy(), L_ref) <= tol_for(np.float32, n)
    vd = downdate_vector(np.tril(L_ref), 27)
    L_dd = np.array(L_ref, order="F")
    oracle.downdate(L_dd, vd.copy())
    rank_1.downdate(Lt, torch.from_numpy(vd.copy()).cuda(), check_diag=True, overwrite_L=True, overwrite_v=True)
    assert rel_fro(Lt.cpu().numpy(), L_dd) <= tol_for(np.float32, n) * 30
