"""Multi-GPU parity + timing of the block-row-cyclic update (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29511 tests/multi_gpu_rowcyclic.py [n_parity] [n_time] [updates] [K_wave]

Parity: every rank builds the same seeded factor, keeps its own rows, applies 3 successive updates
through RowCyclicFactor (NCCL broadcast of each panel's coefficients) and compares its rows with the
oracle (tolerance 1e-12 * sqrt(n)); then K successive updates through update_many (wavefront, one
all-gather per anti-diagonal).  Timing: n_time with per-rank generated rows, both forms.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from cholupdates_b200.distributed import RowCyclicFactor, rc_owned_rows  # noqa: E402
from helpers import synth_factor, tol_for  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    updates = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    k_wave = int(sys.argv[4]) if len(sys.argv) > 4 else 64

    import seeger_oracle as oracle

    L = synth_factor(n, 5)
    rows = rc_owned_rows(n, rank, world)
    A = torch.from_numpy(np.ascontiguousarray(L[rows])).to(dev)
    fac = RowCyclicFactor(A, n, rank, world)
    L_ref = np.array(L, order="F")
    for u in range(3):
        v = np.random.default_rng([1, u]).standard_normal(n)
        fac.update(torch.from_numpy(v.copy()).to(dev), per_panel=(u == 1))  # both single-update forms
        oracle.update(L_ref, v.copy())
    lower = np.arange(n)[None, :] <= rows[:, None]
    got = np.where(lower, A.cpu().numpy(), 0.0)
    want = np.tril(L_ref)[rows]
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    ok = err <= tol_for(np.float64, n) and bool(np.all(A.cpu().numpy()[~lower] == 0.0))
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print(f"[rowcyclic parity] n={n} world={world}: rel. Frobenius error (rank 0 rows) {err:.2e} "
              f"(bar {tol_for(np.float64, n):.1e}); all ranks ok: {int(t.item()) == 0}", flush=True)
    assert int(t.item()) == 0

    # wavefront form: K successive updates in one call
    facw = None
    for K, p2p in ((3, None), (2 * ((n + 255) // 256) + 1, None), (5, False), (4, None)):
        A = torch.from_numpy(np.ascontiguousarray(L[rows])).to(dev)
        V = np.stack([np.random.default_rng([2, u]).standard_normal(n) for u in range(K)])
        if facw is None:
            facw = RowCyclicFactor(A, n, rank, world)  # one factor object: the mailbox epochs carry over
        else:
            facw.A.copy_(A)
            A = facw.A
        facw.update_many(torch.from_numpy(V.copy()).to(dev), p2p=p2p)
        L_ref = np.array(L, order="F")
        for u in range(K):
            oracle.update(L_ref, V[u].copy())
        got = np.where(lower, A.cpu().numpy(), 0.0)
        want = np.tril(L_ref)[rows]
        err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        ok = err <= tol_for(np.float64, n) and bool(np.all(A.cpu().numpy()[~lower] == 0.0))
        t = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(t)
        if rank == 0:
            mode = "NCCL all-gather" if p2p is False else ("peer-memory exchange" if getattr(facw.backend, "_p2p_state", False) else "NCCL all-gather (no peer access)")
            print(f"[rowcyclic wavefront parity] n={n} world={world} K={K} {mode}: rel. Frobenius error (rank 0 rows) "
                  f"{err:.2e} (bar {tol_for(np.float64, n):.1e}); all ranks ok: {int(t.item()) == 0}", flush=True)
        assert int(t.item()) == 0

    # distributed downdate (panel-wise forward substitution + local sweep), valid and not-PD vectors
    from helpers import downdate_vector
    A = torch.from_numpy(np.ascontiguousarray(L[rows])).to(dev)
    vd = downdate_vector(L, 21)
    facd = RowCyclicFactor(A, n, rank, world)
    facd.downdate(torch.from_numpy(vd.copy()).to(dev))                      # wavefront kernels if peer memory
    L_ref = np.array(L, order="F")
    oracle.downdate(L_ref, vd.copy())
    vd2 = downdate_vector(np.tril(L_ref), 23)
    facd.downdate(torch.from_numpy(vd2.copy()).to(dev), per_panel=True)      # per panel + broadcast
    oracle.downdate(L_ref, vd2.copy())
    got = np.where(lower, A.cpu().numpy(), 0.0)
    want = np.tril(L_ref)[rows]
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    keep = A.clone()
    try:
        facd.downdate(torch.from_numpy(downdate_vector(np.tril(L_ref), 22, norm_p=1.2)).to(dev))
        raised = False
    except np.linalg.LinAlgError:
        raised = True
    ok = err <= tol_for(np.float64, n) and raised and bool(torch.equal(keep, A))
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print(f"[rowcyclic downdate parity] n={n} world={world}: rel. Frobenius error (rank 0 rows) {err:.2e} "
              f"(bar {tol_for(np.float64, n):.1e}); not-PD raises with the factor untouched: {raised}; "
              f"all ranks ok: {int(t.item()) == 0}", flush=True)
    assert int(t.item()) == 0

    # timing at n_time, rows generated per rank
    n = n_time
    rows = rc_owned_rows(n, rank, world)
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    A = torch.randn((len(rows), n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
    rt = torch.from_numpy(rows).to(dev)
    A.masked_fill_(torch.arange(n, device=dev)[None, :] > rt[:, None], 0.0)
    A[torch.arange(len(rows), device=dev), rt] = 1.0 + torch.rand(len(rows), dtype=torch.float64, device=dev, generator=g)
    fac = RowCyclicFactor(A, n, rank, world)
    gv = torch.Generator(device=dev)
    gv.manual_seed(7)  # same vectors on every rank
    vs = torch.randn((updates + 1, n), dtype=torch.float64, device=dev, generator=gv)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for per_panel in (True, False):
        fac.update(vs[updates], check_diag=False, per_panel=per_panel)  # warm-up
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for u in range(updates):
            fac.update(vs[u], check_diag=False, per_panel=per_panel)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            per = float(ms.item()) / updates
            bytes_upd = 8 * n * (n + 1) + 16 * n
            print(f"[rowcyclic timing] n={n} world={world} single update, "
                  f"{'launch-per-panel + broadcast' if per_panel else 'K=1 wavefront'}: "
                  f"{per:.3f} ms/update, {bytes_upd / per / 1e6:.0f} GB/s aggregate "
                  f"({bytes_upd / per / 1e6 / world:.0f} GB/s per GPU)", flush=True)
    # one downdate (v = 0.5 * L u / |u| is always valid)
    for per_panel in (True, None):
        uu = torch.randn(n, dtype=torch.float64, device=dev, generator=gv)
        uu = 0.5 * uu / uu.norm()
        vloc = torch.zeros(n, dtype=torch.float64, device=dev)
        vloc[rt] = A @ uu  # the rows of L u this rank owns
        dist.all_reduce(vloc)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        fac.downdate(vloc, check_diag=False, per_panel=per_panel)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            how = "forward substitution per panel + broadcast" if per_panel else "forward substitution on the wavefront kernels"
            print(f"[rowcyclic downdate timing] n={n} world={world} {how}: {float(ms.item()):.3f} ms per downdate", flush=True)

    # wavefront: k_wave successive updates in one call (BASELINE.json configs[4]: 64)
    gv.manual_seed(8)
    V = torch.randn((k_wave, n), dtype=torch.float64, device=dev, generator=gv)
    for p2p in (False, None):
        fac.update_many(V[: min(4, k_wave)], check_diag=False, p2p=p2p)  # warm-up
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        fac.update_many(V, check_diag=False, p2p=p2p)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            per = float(ms.item()) / k_wave
            bytes_upd = 8 * n * (n + 1) + 16 * n
            mode = "NCCL all-gather" if p2p is False else ("peer-memory exchange" if getattr(fac.backend, "_p2p_state", False) else "NCCL all-gather (no peer access)")
            print(f"[rowcyclic wavefront timing] n={n} world={world} K={k_wave} {mode}: {per:.3f} ms/update "
                  f"({float(ms.item()):.1f} ms total), {bytes_upd / per / 1e6:.0f} GB/s aggregate "
                  f"({bytes_upd / per / 1e6 / world:.0f} GB/s per GPU)", flush=True)
    finite = bool(torch.isfinite(A).all().item())
    assert finite
    fac.close()  # unmap / free the peer-memory mailboxes (collective)
    facw.close()
    facd.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
