"""world_size-2 gloo test (CPU) of the batched multi-GPU mode's host logic: balanced sharding by rank,
local processing of the shard, all-gather of the per-factor status vector, identical exception on
every rank.  The local kernel call is replaced by the oracle (no GPU here); on the B200 box the same
code path runs with the CUDA kernels (tests/test_gpu_parity.py covers those)."""

import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from cholupdates_b200.distributed import shard_bounds  # noqa: E402


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _oracle_batched(op):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import seeger_oracle as oracle

    def fn(L, v, check_diag=True, overwrite_L=False, overwrite_v=False, raise_on_error=False):
        Ln, vn = L.numpy().copy(), v.numpy().copy()
        status = np.zeros(Ln.shape[0], dtype=np.int32)
        for b in range(Ln.shape[0]):
            if check_diag and np.any(np.diag(Ln[b]) == 0):
                status[b] = 1
                continue
            try:
                Lb = np.ascontiguousarray(Ln[b])
                getattr(oracle, op)(Lb, vn[b])
                Ln[b] = Lb
            except np.linalg.LinAlgError:
                status[b] = 2
                Ln[b] = L.numpy()[b]
        return torch.from_numpy(Ln), torch.from_numpy(status)

    return fn


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from cholupdates_b200.distributed import shard_bounds, sharded_batched
        from helpers import downdate_vector, synth_factor

        B, n = 7, 24  # uneven split: 4 + 3
        Ls = np.stack([synth_factor(n, [0, b]) for b in range(B)])
        vs = np.stack([downdate_vector(Ls[b], [1, b]) for b in range(B)])
        vs[5] = downdate_vector(Ls[5], 3, norm_p=1.3)  # lives on rank 1: not positive definite
        Ls[1, 2, 2] = 0.0                               # lives on rank 0: zero on the diagonal
        lo, hi = shard_bounds(B, rank, world)
        L_loc, v_loc = torch.from_numpy(Ls[lo:hi].copy()), torch.from_numpy(vs[lo:hi].copy())
        out, status = sharded_batched("downdate", L_loc, v_loc, raise_on_error=False,
                                      local_fn=_oracle_batched("downdate"))
        assert status.tolist() == [0, 1, 0, 0, 0, 2, 0], status.tolist()
        # local results: failing factors untouched, the others changed
        for b in range(lo, hi):
            same = np.array_equal(out[b - lo].numpy(), Ls[b])
            assert same == (b in (1, 5)), b
        # every rank raises the same error
        try:
            sharded_batched("downdate", L_loc, v_loc, local_fn=_oracle_batched("downdate"))
            raised = ""
        except np.linalg.LinAlgError as exc:
            raised = str(exc)
        assert "factor 1 of the global batch; 2 failed" in raised, raised
        with open(os.path.join(tmpdir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_batched_world_size_2_gloo(tmp_path):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


# ---------------------------------------------------------------- block-row-cyclic single factor
def _rc_worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import seeger_oracle as oracle
        from cholupdates_b200.distributed import RowCyclicFactor, rc_owned_rows
        from helpers import rel_fro, synth_factor, tol_for
        from rc_numpy_backend import NumpyRowCyclicBackend

        n = 900  # 4 blocks of 256 rows (the last one partial): rank 0 owns blocks 0, 2; rank 1 owns 1, 3
        L = synth_factor(n, 3)
        L[:, 5] *= -1.0  # one negative diagonal entry: the sign fix must survive the distribution
        rows = rc_owned_rows(n, rank, world)
        A_loc = torch.from_numpy(np.ascontiguousarray(L[rows]))
        fac = RowCyclicFactor(A_loc, n, rank, world, backend=NumpyRowCyclicBackend(A_loc, n, rank, world))
        L_ref = np.array(L, order="F")
        for u in range(3):  # successive updates, as in BASELINE.json configs[4]
            v = np.random.default_rng([1, u]).standard_normal(n)
            fac.update(torch.from_numpy(v.copy()), per_panel=(u == 1))  # both single-update forms
            oracle.update(L_ref, v.copy())
        lower = np.arange(n)[None, :] <= rows[:, None]  # lower triangle in GLOBAL row indices
        got = np.where(lower, A_loc.numpy(), 0.0)
        want = np.tril(L_ref)[rows]
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err <= tol_for(np.float64, n), err
        assert np.all(A_loc.numpy()[~lower] == 0.0)  # nothing above the diagonal was written
        # K successive updates as a wavefront (one all-gather per anti-diagonal of (update, panel) tasks):
        # same result as K calls of update(); K > number of panels and K < number of panels both occur
        for K in (2, 7):
            A2 = torch.from_numpy(np.ascontiguousarray(L[rows]))
            fac2 = RowCyclicFactor(A2, n, rank, world, backend=NumpyRowCyclicBackend(A2, n, rank, world))
            V = np.stack([np.random.default_rng([2, u]).standard_normal(n) for u in range(K)])
            z = fac2.update_many(torch.from_numpy(V.copy()))
            assert z.shape == (K, n) and bool(torch.all(z == 0.5))  # every rank sees every panel's scratch
            L_ref2 = np.array(L, order="F")
            for u in range(K):
                oracle.update(L_ref2, V[u].copy())
            got2 = np.where(lower, A2.numpy(), 0.0)
            want2 = np.tril(L_ref2)[rows]
            err2 = np.linalg.norm(got2 - want2) / np.linalg.norm(want2)
            assert err2 <= tol_for(np.float64, n), (K, err2)
            assert np.all(A2.numpy()[~lower] == 0.0)
        # distributed downdate: panel-wise forward substitution + local sweep == the oracle's downdate
        from helpers import downdate_vector
        A3 = torch.from_numpy(np.ascontiguousarray(L[rows]))
        fac3 = RowCyclicFactor(A3, n, rank, world, backend=NumpyRowCyclicBackend(A3, n, rank, world))
        vd = downdate_vector(L, 11)
        fac3.downdate(torch.from_numpy(vd.copy()))
        L_ref3 = np.array(L, order="F")
        oracle.downdate(L_ref3, vd.copy())
        got3 = np.where(lower, A3.numpy(), 0.0)
        want3 = np.tril(L_ref3)[rows]
        err3 = np.linalg.norm(got3 - want3) / np.linalg.norm(want3)
        assert err3 <= tol_for(np.float64, n), err3
        assert np.all(A3.numpy()[~lower] == 0.0)
        # not positive definite: every rank raises, the factor is untouched
        keep3 = A3.clone()
        try:
            fac3.downdate(torch.from_numpy(downdate_vector(np.tril(L_ref3), 12, norm_p=1.2)))
            raised = False
        except np.linalg.LinAlgError as exc:
            raised = "not positive definite" in str(exc)
        assert raised and torch.equal(keep3, A3)
        # zero on the diagonal of a row owned by rank 1: both ranks raise, nothing is modified
        before = A_loc.clone()
        if 300 in rows:
            A_loc[int(np.searchsorted(rows, 300)), 300] = 0.0
            before = A_loc.clone()
        try:
            fac.update(torch.from_numpy(np.ones(n)))
            raised = False
        except np.linalg.LinAlgError:
            raised = True
        assert raised and torch.equal(before, A_loc)
        try:
            fac.update_many(torch.from_numpy(np.ones((3, n))))
            raised = False
        except np.linalg.LinAlgError:
            raised = True
        assert raised and torch.equal(before, A_loc)
        with open(os.path.join(tmpdir, f"rc{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_rowcyclic_update_world_size_2_gloo(tmp_path):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_rc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "rc0").exists() and (tmp_path / "rc1").exists()


def test_rc_owned_rows_partition():
    from cholupdates_b200.distributed import rc_owned_rows

    for n in (1, 255, 256, 257, 900, 4096):
        for world in (1, 2, 3, 8):
            allrows = np.concatenate([rc_owned_rows(n, r, world) for r in range(world)])
            assert sorted(allrows.tolist()) == list(range(n))
