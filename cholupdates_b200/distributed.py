"""Multi-GPU mode for batches of independent factors (BASELINE.json configs[3]).

Independent factors shard naturally: each rank (one process per GPU) owns a contiguous slice of the
batch, runs the batched kernel on it, and nothing crosses NVLink on the data path.  The only exchange
is the optional all-gather of the per-factor status vector, so that every rank can raise the same
``LinAlgError`` -- a few KB over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

The single large factor does not shard on one node's worth of memory pressure (n = 16384 fits one
GPU): ``bench.py --gpus N`` runs replicas.  The block-row-cyclic mode for one huge factor
(``RowCyclicFactor``) is below.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice ``[lo, hi)`` of a batch of ``batch`` factors owned by ``rank``
    (the first ``batch % world_size`` ranks own one factor more)."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    q, r = divmod(max(int(batch), 0), world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def _default_local(op: str) -> Callable:
    from . import rank_1

    return rank_1.update_batched if op == "update" else rank_1.downdate_batched


def sharded_batched(
    op: str,
    L_local,
    v_local,
    check_diag: bool = True,
    overwrite_L: bool = False,
    overwrite_v: bool = False,
    raise_on_error: bool = True,
    group=None,
    local_fn: Optional[Callable] = None,
):
    """Run ``update_batched`` / ``downdate_batched`` on this rank's shard and gather the status
    vectors of all ranks.

    ``L_local``/``v_local`` are the rank's own slice (``shard_bounds``).  Returns
    ``(L_local_out, status_global)`` where ``status_global`` concatenates the shards in rank order
    (shards may differ in length by one).  With ``raise_on_error`` every rank raises the same
    ``numpy.linalg.LinAlgError`` if any factor anywhere failed; the failing factors are untouched.
    ``local_fn`` replaces the local kernel call (the CPU tests inject the oracle here).
    """
    import numpy as np
    import torch
    import torch.distributed as dist

    fn = local_fn or _default_local(op)
    L_out, status = fn(L_local, v_local, check_diag=check_diag, overwrite_L=overwrite_L,
                       overwrite_v=overwrite_v, raise_on_error=False)
    status = status.to(torch.int32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        count = torch.tensor([status.numel()], dtype=torch.int64, device=status.device)
        counts = [torch.zeros_like(count) for _ in range(world)]
        dist.all_gather(counts, count, group=group)
        width = int(max(int(c.item()) for c in counts))
        padded = torch.zeros(max(width, 1), dtype=torch.int32, device=status.device)
        padded[: status.numel()] = status
        gathered = [torch.zeros_like(padded) for _ in range(world)]
        dist.all_gather(gathered, padded, group=group)
        status_global = torch.cat([g[: int(c.item())] for g, c in zip(gathered, counts)])
    else:
        status_global = status
    if raise_on_error:
        bad = torch.nonzero(status_global).flatten()
        if bad.numel() > 0:
            b = int(bad[0].item())
            code = int(status_global[b].item())
            what = ("The given Cholesky factor `L` contains zeros on its diagonal" if code == 1
                    else "The downdated matrix is not positive definite.")
            raise np.linalg.LinAlgError(f"{what} (factor {b} of the global batch; {bad.numel()} failed)")
    return L_out, status_global


# =====================================================================================
# One huge factor, block-row-cyclic over the ranks (BASELINE.json configs[4])
# =====================================================================================
RC_BLOCK = 256  # rows per distribution block = columns per panel (csrc/rowcyclic.cuh: kRcBlock)


def rc_owned_rows(n: int, rank: int, world_size: int):
    """Global row indices owned by ``rank`` (blocks of 256 rows dealt cyclically), in local order."""
    import numpy as np

    rows = []
    for blk in range(rank, (n + RC_BLOCK - 1) // RC_BLOCK, world_size):
        rows.append(np.arange(blk * RC_BLOCK, min(n, (blk + 1) * RC_BLOCK)))
    return np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)


class _CudaRowCyclicBackend:
    """Per-panel pieces through the C ABI (``chub_rc_*``, include/cholupdates_b200.h)."""

    def __init__(self, A_local, n, rank, world):
        import torch

        from . import _lib

        self.lib = _lib.require_device()
        self.torch = torch
        self.A, self.n, self.rank, self.world = A_local, n, rank, world
        self.suf = "f64" if A_local.dtype == torch.float64 else "f32"
        dev = A_local.device
        self.ws = torch.empty(int(self.lib.chub_workspace_bytes(1, n)), dtype=torch.uint8, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.payload_doubles = int(self.lib.chub_rc_payload_doubles())
        import ctypes

        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        for suf in ("f64", "f32"):
            f = getattr(self.lib, f"chub_rc_prepare_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, vp, vp, ci, vp, vp]
            f = getattr(self.lib, f"chub_rc_panel_owner_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, ci, ci, ci, vp, vp, vp, vp, vp]
            f = getattr(self.lib, f"chub_rc_panel_apply_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, ci, ci, vp, vp, vp, vp]

    def _stream(self):
        return self.torch.cuda.current_stream(self.A.device).cuda_stream

    def new_payload(self):
        return self.torch.empty(self.payload_doubles, dtype=self.torch.float64, device=self.A.device)

    def prepare(self, v, v_out, check_diag):
        from . import _lib

        self.v_out = v_out
        rc = getattr(self.lib, f"chub_rc_prepare_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank,
            v.data_ptr(), self.ws.data_ptr(), _lib.FLAG_CHECK_DIAG if check_diag else 0,
            self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rc_prepare")

    def panel_owner(self, J, payload):
        from . import _lib

        rc = getattr(self.lib, f"chub_rc_panel_owner_{self.suf}")(
            self.A.data_ptr(), self.n, self.A.shape[1], self.world, self.rank, J, self.v_out.data_ptr(),
            self.ws.data_ptr(), payload.data_ptr(), self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rc_panel_owner")

    def panel_apply(self, J, is_owner, payload):
        from . import _lib

        rc = getattr(self.lib, f"chub_rc_panel_apply_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank, J,
            1 if is_owner else 0, self.ws.data_ptr(), payload.data_ptr(), self.status.data_ptr(),
            self._stream())
        _lib.check(rc, "chub_rc_panel_apply")

    def status_tensor(self):
        return self.status

    # ---- downdate (chub_rc_dd_*)
    def _bind_dd(self):
        import ctypes

        if getattr(self, "_dd_bound", False):
            return
        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        for suf in ("f64", "f32"):
            f = getattr(self.lib, f"chub_rc_dd_panel_owner_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, ci, ci, ci, vp, vp, vp, vp]
            f = getattr(self.lib, f"chub_rc_dd_panel_apply_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, ci, vp, vp, vp, vp]
            f = getattr(self.lib, f"chub_rc_dd_finish_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, vp, vp, vp, vp]
        self._dd_bound = True

    def dd_panel_owner(self, J, payload):
        from . import _lib

        self._bind_dd()
        rc = getattr(self.lib, f"chub_rc_dd_panel_owner_{self.suf}")(
            self.A.data_ptr(), self.n, self.A.shape[1], self.world, self.rank, J, self.ws.data_ptr(),
            payload.data_ptr(), self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rc_dd_panel_owner")

    def dd_panel_apply(self, J, payload):
        from . import _lib

        self._bind_dd()
        rc = getattr(self.lib, f"chub_rc_dd_panel_apply_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank, J,
            self.ws.data_ptr(), payload.data_ptr(), self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rc_dd_panel_apply")

    def dd_finish(self):
        """rho^2 test + coefficients (every rank, identical) and the local sweep; returns the status."""
        from . import _lib

        self._bind_dd()
        rc = getattr(self.lib, f"chub_rc_dd_finish_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank,
            self.v_out.data_ptr(), self.ws.data_ptr(), self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rc_dd_finish")
        return int(self.status.item())

    # ---- wavefront over K successive updates (chub_rcw_*)
    def wave_begin(self, V, V_out, check_diag):
        """Allocate the step buffers, W[u] <- V[u] on the owned rows; returns (pay_mine, pay_all)."""
        import ctypes

        from . import _lib

        torch, lib = self.torch, self.lib
        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        lib.chub_rcw_workspace_bytes.restype = ctypes.c_size_t
        lib.chub_rcw_workspace_bytes.argtypes = [i64, i64, ci]
        lib.chub_rcw_max_tasks.argtypes = [i64, i64, ci]
        for suf in ("f64", "f32"):
            f = getattr(lib, f"chub_rcw_prepare_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, vp, i64, i64, vp, ci, vp, vp]
            f = getattr(lib, f"chub_rcw_owner_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, ci, ci, i64, i64, vp, vp, ci, ci, vp, vp]
            f = getattr(lib, f"chub_rcw_apply_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, i64, i64, vp, vp, vp, i64, ci, vp, ci, vp, vp]
        K, dev = int(V.shape[0]), self.A.device
        self.K, self.V_out = K, V_out
        self.fuse = 1
        need = int(lib.chub_rcw_workspace_bytes(self.n, K, self.world))
        if getattr(self, "wave_ws", None) is None or self.wave_ws.numel() < need:
            self.wave_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        slots = int(lib.chub_rcw_max_tasks(self.n, K, self.world)) * int(lib.chub_rcw_payload_doubles())
        pay_all = torch.empty(self.world * slots, dtype=torch.float64, device=dev)
        pay_mine = pay_all[self.rank * slots:(self.rank + 1) * slots] if self.world == 1 else \
            torch.empty(slots, dtype=torch.float64, device=dev)
        rc = getattr(lib, f"chub_rcw_prepare_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank,
            V.data_ptr(), K, V.stride(0), self.wave_ws.data_ptr(),
            _lib.FLAG_CHECK_DIAG if check_diag else 0, self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rcw_prepare")
        return pay_mine, pay_all

    # ---- exchange over NVLink peer memory (cudaIpc mailboxes) instead of NCCL
    def p2p_ready(self, group=None) -> bool:
        """Set up the mailboxes once (collective: every rank must call it); True if every rank could map
        every peer's mailbox.  CHUB_RC_P2P=0 disables the peer-memory exchange."""
        import ctypes
        import os

        import torch.distributed as dist

        if getattr(self, "_p2p_state", None) is not None:
            return self._p2p_state
        torch, lib = self.torch, self.lib
        vp, i64, ci, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
        lib.chub_rcw_mailbox_bytes.restype, lib.chub_rcw_mailbox_bytes.argtypes = sz, [i64, i64, ci]
        lib.chub_p2p_alloc.restype, lib.chub_p2p_alloc.argtypes = ci, [sz, ctypes.POINTER(vp)]
        lib.chub_p2p_export.restype, lib.chub_p2p_export.argtypes = ci, [vp, vp]
        lib.chub_p2p_import.restype, lib.chub_p2p_import.argtypes = ci, [vp, ctypes.POINTER(vp)]
        for suf in ("f64", "f32"):
            f = getattr(lib, f"chub_rcw_owner_p2p_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, ci, ci, i64, i64, vp, vp, vp, vp, i64, ci, ci, vp, vp]
            f = getattr(lib, f"chub_rcw_apply_p2p_{suf}")
            f.restype, f.argtypes = ci, [vp, i64, i64, i64, ci, ci, i64, i64, vp, vp, vp, vp, i64, i64, ci, vp, ci, vp, vp]
        ok = os.environ.get("CHUB_RC_P2P", "1") != "0"
        ptrs = [0] * self.world
        if ok:
            n_panels = (self.n + RC_BLOCK - 1) // RC_BLOCK
            nbytes = int(lib.chub_rcw_mailbox_bytes(self.n, n_panels, self.world))  # enough for every K
            mine = vp()
            ok = lib.chub_p2p_alloc(nbytes, ctypes.byref(mine)) == 0
            handle = (ctypes.c_ubyte * 64)()
            ok = ok and lib.chub_p2p_export(mine, handle) == 0
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
        if ok and all(h is not None for h in handles):
            ptrs[self.rank] = mine.value
            for g, h in enumerate(handles):
                if g == self.rank:
                    continue
                peer = vp()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                if lib.chub_p2p_import(buf, ctypes.byref(peer)) != 0:
                    ok = False
                    break
                ptrs[g] = peer.value
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.A.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        ok = bool(int(flag.item()))
        if ok:
            self.mailbox = mine.value
            self.peer_table = torch.tensor(ptrs, dtype=torch.int64, device=self.A.device)
            self.epoch = 0
        self._p2p_state = ok
        return ok

    def p2p_close(self, group=None):
        """Collective: unmap the peers' mailboxes and free this rank's (after every rank has stopped using them)."""
        import ctypes

        import torch.distributed as dist

        if not getattr(self, "_p2p_state", False):
            return
        self.torch.cuda.synchronize(self.A.device)
        dist.barrier(group=group)
        self.lib.chub_p2p_close.restype, self.lib.chub_p2p_close.argtypes = ctypes.c_int, [ctypes.c_void_p]
        self.lib.chub_p2p_free.restype, self.lib.chub_p2p_free.argtypes = ctypes.c_int, [ctypes.c_void_p]
        for g, ptr in enumerate(self.peer_table.tolist()):
            if g != self.rank:
                self.lib.chub_p2p_close(ptr)
        dist.barrier(group=group)
        self.lib.chub_p2p_free(self.mailbox)
        self._p2p_state = None
        self.mailbox = None

    def wave_owner_p2p(self, d, pay_mine, solve_only=False):
        from . import _lib

        rc = getattr(self.lib, f"chub_rcw_owner_p2p_{self.suf}")(
            self.A.data_ptr(), self.n, self.A.shape[1], self.world, self.rank, self.K, d,
            self.wave_ws.data_ptr(), pay_mine.data_ptr(), self.peer_table.data_ptr(), self.mailbox, self.epoch,
            1 if solve_only else 0, 1 if solve_only else self.fuse, self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rcw_owner_p2p")

    def wave_apply_p2p(self, d, rows_mode=0, solve_only=False):
        from . import _lib

        rc = getattr(self.lib, f"chub_rcw_apply_p2p_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank, self.K, d,
            self.wave_ws.data_ptr(), self.peer_table.data_ptr(), self.mailbox, self.V_out.data_ptr(),
            self.V_out.stride(0), self.epoch, rows_mode, self.ws.data_ptr() if solve_only else None,
            1 if solve_only else self.fuse, self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rcw_apply_p2p")

    def lookahead_streams(self):
        """(head, bulk): a high-priority stream for owner + head kernels and one for the bulk of the trail."""
        if getattr(self, "_la_streams", None) is None:
            dev = self.A.device
            self._la_streams = (self.torch.cuda.Stream(device=dev, priority=-1), self.torch.cuda.Stream(device=dev))
        return self._la_streams

    def wave_end_p2p(self, steps):
        """Advance the epoch; one status read (a peer that never arrived trips the device watchdog)."""
        self.epoch += steps
        if int(self.status.item()) == 3:
            # Nothing was stored with a stale payload (the trail CTAs bail out), but the panels already
            # applied stay applied: the factor is in an undefined intermediate state on every rank, and
            # the flags no longer match the epoch -- the mailbox must be re-created (a new RowCyclicFactor).
            self._p2p_state = False
            raise RuntimeError("cholupdates_b200: a peer did not deliver its panel coefficients within "
                               "CHUB_P2P_WATCHDOG_MS (device watchdog, CHUB_STATUS_INTERNAL); the factor "
                               "is invalid and the peer-memory exchange of this object is disabled")

    def wave_owner(self, d, pay_mine, solve_only=False):
        from . import _lib

        rc = getattr(self.lib, f"chub_rcw_owner_{self.suf}")(
            self.A.data_ptr(), self.n, self.A.shape[1], self.world, self.rank, self.K, d,
            self.wave_ws.data_ptr(), pay_mine.data_ptr(), 1 if solve_only else 0, 1 if solve_only else self.fuse,
            self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rcw_owner")

    def wave_apply(self, d, pay_all, rows_mode=0, solve_only=False):
        from . import _lib

        rc = getattr(self.lib, f"chub_rcw_apply_{self.suf}")(
            self.A.data_ptr(), self.A.shape[0], self.n, self.A.shape[1], self.world, self.rank, self.K, d,
            self.wave_ws.data_ptr(), pay_all.data_ptr(), self.V_out.data_ptr(), self.V_out.stride(0),
            rows_mode, self.ws.data_ptr() if solve_only else None, 1 if solve_only else self.fuse,
            self.status.data_ptr(), self._stream())
        _lib.check(rc, "chub_rcw_apply")

    supports_wave_solve = True  # the forward substitution of the downdate can run on the wavefront kernels


def _check_vector_like(A, x, what: str) -> None:
    """dtype / device agreement between the local rows and a vector argument (the reference's dtype
    checks, _seeger.py:243-258, for the distributed entry points)."""
    import torch

    if A.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"The local rows must be single or double precision (given: {A.dtype}).")
    if x.dtype != A.dtype:
        raise TypeError(f"`{what}` must have the dtype of the factor ({A.dtype}), but got {x.dtype}.")
    if x.device != A.device:
        raise ValueError(f"`{what}` must live on the device of the factor ({A.device}), but is on {x.device}.")


class RowCyclicFactor:
    """A Cholesky factor of order ``n`` distributed block-row-cyclically: this rank holds
    ``A_local`` = the rows ``rc_owned_rows(n, rank, world)`` (all ``n`` columns, row-major).

    ``update(v)`` applies the rank-1 Seeger update ``L L^T + v v^T`` in place on every rank's rows.
    Per 256-column panel the owner of the diagonal block (rank ``J mod world``) computes the panel's
    coefficients and broadcasts ~8 KB; everything else is local.  ``v`` is the full
    vector, identical on every rank.
    """

    def __init__(self, A_local, n: int, rank: int = 0, world_size: int = 1, group=None, backend=None):
        self.A = A_local
        self.n, self.rank, self.world, self.group = int(n), int(rank), int(world_size), group
        expect = len(rc_owned_rows(self.n, self.rank, self.world))
        if A_local.ndim != 2 or A_local.shape[0] != expect or A_local.shape[1] < self.n:
            raise ValueError(
                f"`A_local` must have shape ({expect}, >= {self.n}) on rank {rank} of {world_size} "
                f"(given shape: {tuple(A_local.shape)}).")
        if not A_local.is_contiguous():
            raise ValueError("`A_local` must be row-major contiguous.")
        self.backend = backend if backend is not None else _CudaRowCyclicBackend(A_local, self.n, rank, world_size)

    def close(self):
        """Release the peer-memory mailboxes (collective over the group; the factor's rows are untouched).
        Optional: a process that keeps one factor for its lifetime can skip it."""
        if hasattr(self.backend, "p2p_close"):
            self.backend.p2p_close(self.group)

    def update(self, v, check_diag: bool = True, per_panel=None):
        """In-place update of the distributed factor.  Returns the scratch vector (rotg's z_k; with
        ``per_panel`` only for the panels this rank owned).  Raises ``numpy.linalg.LinAlgError`` on
        every rank if ``check_diag`` finds a zero on the diagonal anywhere (nothing is modified).

        ``per_panel=True``: per 256-column panel the owner's kernels, one broadcast, the apply kernels
        (``chub_rc_panel_owner`` / ``chub_rc_panel_apply``).  ``per_panel=False``: the K = 1 case of
        ``update_many`` (one fused owner kernel, one all-gather, one trail kernel per panel).  Both are
        bound by per-panel launch / exchange latency.  With NCCL the broadcast form is the faster one
        (8 B200, n = 65536: 27 ms vs 64 ms -- ``all_gather_into_tensor`` costs more host time per call
        than ``broadcast``); when the ranks can map each other's mailboxes (``update_many``'s
        peer-memory exchange) the wavefront form needs no collective at all and is the default.
        Successive vectors belong in ``update_many``."""
        import numpy as np
        import torch
        import torch.distributed as dist

        if v.ndim != 1 or v.shape[0] != self.n:
            raise ValueError(
                f"The shape of the given vector `v` is incompatible with the shape of the given "
                f"Cholesky factor `L`. Expected shape {(self.n,)} but got {tuple(v.shape)}.")
        _check_vector_like(self.A, v, "v")
        if not v.is_contiguous():
            raise ValueError("`v` must be contiguous.")
        if per_panel is None:
            # with an exchange: the wavefront code only if it can use peer memory (no collective per panel)
            per_panel = self.world > 1 and not (hasattr(self.backend, "p2p_ready")
                                                and self.backend.p2p_ready(self.group))
        if not per_panel and hasattr(self.backend, "wave_begin"):
            return self.update_many(v.reshape(1, -1), check_diag=check_diag)[0]
        be = self.backend
        multi = self.world > 1 and dist.is_available() and dist.is_initialized()
        v_out = torch.zeros_like(v)
        be.prepare(v, v_out, check_diag)
        if check_diag:
            st = be.status_tensor()
            if multi:
                dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
            if int(st.item()) == 1:
                raise np.linalg.LinAlgError("The given Cholesky factor `L` contains zeros on its diagonal")
        payload = be.new_payload()
        for J in range((self.n + RC_BLOCK - 1) // RC_BLOCK):
            owner = J % self.world
            if owner == self.rank:
                be.panel_owner(J, payload)
            if multi:
                src = owner if self.group is None else dist.get_global_rank(self.group, owner)
                dist.broadcast(payload, src=src, group=self.group)
            be.panel_apply(J, owner == self.rank, payload)
        return v_out

    def downdate(self, v, check_diag: bool = True, per_panel=None):
        """In-place rank-1 downdate ``L L^T - v v^T`` of the distributed factor (reference semantics:
        ``_seeger_impl_cython.pyx:160-330``).  The forward substitution ``p = L^{-1} v`` runs panel by
        panel (the owner of each diagonal block broadcasts the panel's ``p``; every rank updates the
        residual of its rows below); ``rho^2 = 1 - p.p`` is then formed identically on every rank, and
        the reverse sweep touches only local rows -- no further exchange.  Returns the
        scratch vector (rotg's z_k; ``p`` if the result is not positive definite).  Raises
        ``numpy.linalg.LinAlgError`` on every rank, with the factor untouched, for a zero on the
        diagonal (``check_diag``) or when the downdated matrix is not positive definite.

        ``per_panel``: None = the forward substitution runs on the wavefront kernels (one owner kernel
        and one read-only trail kernel per panel) when their exchange needs no collective (peer memory,
        or a single rank), else per panel with an NCCL broadcast; True = always the latter."""
        import numpy as np
        import torch
        import torch.distributed as dist

        if v.ndim != 1 or v.shape[0] != self.n:
            raise ValueError(
                f"The shape of the given vector `v` is incompatible with the shape of the given "
                f"Cholesky factor `L`. Expected shape {(self.n,)} but got {tuple(v.shape)}.")
        _check_vector_like(self.A, v, "v")
        if not v.is_contiguous():
            raise ValueError("`v` must be contiguous.")
        be = self.backend
        multi = self.world > 1 and dist.is_available() and dist.is_initialized()
        v_out = torch.zeros_like(v)
        use_p2p = (multi and per_panel is not True and getattr(be, "supports_wave_solve", False)
                   and be.p2p_ready(self.group))
        if getattr(be, "supports_wave_solve", False) and per_panel is not True and (use_p2p or not multi):
            # forward substitution on the wavefront kernels (K = 1, solve only): one owner kernel and one
            # read-only trail kernel per panel, the panel's p exchanged over peer memory -- no collective
            pay_mine, pay_all = be.wave_begin(v.reshape(1, -1), v_out.reshape(1, -1), check_diag)
            be.v_out = v_out
            if check_diag or use_p2p:  # (peer memory: also the rendezvous of the call, see update_many)
                st = be.status_tensor()
                if multi:
                    dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
                if check_diag and int(st.item()) == 1:
                    raise np.linalg.LinAlgError("The given Cholesky factor `L` contains zeros on its diagonal")
            steps = (self.n + RC_BLOCK - 1) // RC_BLOCK
            for d in range(steps):
                if use_p2p:
                    be.wave_owner_p2p(d, pay_mine, solve_only=True)
                    be.wave_apply_p2p(d, 0, solve_only=True)
                else:
                    be.wave_owner(d, pay_mine, solve_only=True)
                    be.wave_apply(d, pay_all, 0, solve_only=True)
            if use_p2p:
                be.wave_end_p2p(steps)
            if be.dd_finish() == 2:
                raise np.linalg.LinAlgError("The downdated matrix is not positive definite.")
            return v_out
        be.prepare(v, v_out, check_diag)
        if check_diag:
            st = be.status_tensor()
            if multi:
                dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
            if int(st.item()) == 1:
                raise np.linalg.LinAlgError("The given Cholesky factor `L` contains zeros on its diagonal")
        payload = be.new_payload()
        for J in range((self.n + RC_BLOCK - 1) // RC_BLOCK):
            owner = J % self.world
            if owner == self.rank:
                be.dd_panel_owner(J, payload)
            if multi:
                src = owner if self.group is None else dist.get_global_rank(self.group, owner)
                dist.broadcast(payload, src=src, group=self.group)
            be.dd_panel_apply(J, payload)
        if be.dd_finish() == 2:
            raise np.linalg.LinAlgError("The downdated matrix is not positive definite.")
        return v_out

    def update_many(self, V, check_diag: bool = True, p2p=None, lookahead=None, fuse=None):
        """Apply the K successive rank-1 updates ``V[0], V[1], ...`` (``V``: ``(K, n)``, identical on
        every rank) in place -- the same result as K calls of ``update``, scheduled as a wavefront.

        Panel J of update u needs only panel J of update u-1 and panel J-1 of update u, so all
        (u, J) with u + J = d are independent.  Step d: every rank computes the coefficients of the
        diagonal blocks it owns, ONE all-gather ships all of the step's coefficient sets, every rank
        applies them to its local rows.  ``K + ceil(n/256) - 1`` exchanges replace ``K * ceil(n/256)``
        broadcasts.  Returns the ``(K, n)`` scratch vectors (rotg's z_k, complete on
        every rank).  Raises ``LinAlgError`` on every rank, with nothing modified, if ``check_diag``
        finds a zero on the input diagonal.

        ``p2p``: None = exchange over NVLink peer memory when every rank can map every peer's mailbox
        (one node, cudaIpc), else NCCL all-gather; False = NCCL; True = peer memory or raise.
        ``lookahead``: None = split every step's trail into head and bulk on two streams when there is
        no collective in the loop (peer-memory exchange or a single rank) and K > 1; False = one stream.
        ``fuse``: 1 = one update per task; 2 = consecutive updates (2u, 2u+1) share one pass over every
        panel (half the HBM bytes per update; the owner runs the two diagonal-block chains back to back
        on a scratch copy of its block).  Same results.  None (default) = 2 when this rank holds at least
        24576 rows (the trail then dominates the step and halving its bytes pays), else 1."""
        import numpy as np
        import torch
        import torch.distributed as dist

        if V.ndim != 2 or V.shape[1] != self.n:
            raise ValueError(
                f"`V` must hold one vector per row, expected shape (K, {self.n}) but got {tuple(V.shape)}.")
        _check_vector_like(self.A, V, "V")
        K = int(V.shape[0])
        V_out = torch.zeros_like(V, memory_format=torch.contiguous_format)
        if K == 0:
            return V_out
        if V.stride(1) != 1:
            raise ValueError("The rows of `V` must be contiguous.")
        be = self.backend
        multi = self.world > 1 and dist.is_available() and dist.is_initialized()
        use_p2p = multi and p2p is not False and hasattr(be, "p2p_ready") and be.p2p_ready(self.group)
        pay_mine, pay_all = be.wave_begin(V, V_out, check_diag)
        if check_diag or use_p2p:
            # With the peer-memory exchange this all-reduce is also the RENDEZVOUS of the call: it runs
            # on the stream, so a rank whose host arrives seconds late holds the others inside the
            # collective (no watchdog there) instead of inside a kernel that spins on its flags.
            st = be.status_tensor()
            if multi:
                dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
            if check_diag and int(st.item()) == 1:
                raise np.linalg.LinAlgError("The given Cholesky factor `L` contains zeros on its diagonal")
        if lookahead is None:
            lookahead = K > 1
        lookahead = bool(lookahead) and (use_p2p or not multi) and hasattr(be, "lookahead_streams")
        if fuse is None:
            # Two updates per pass halve the bytes the trail moves but double the owner kernel's serial
            # work per step.  Measured (B200): with >= 32768 local rows the trail dominates and fuse = 2 wins
            # (n = 32768 on one GPU: 1.19 vs 1.75 ms per update; n = 65536 on two: 2.27 vs 3.53); with
            # 16384 local rows the owner's latency dominates and it loses (0.69 vs 0.54).
            fuse = 2 if (K >= 2 and int(self.A.shape[0]) >= 24576 and hasattr(be, "lookahead_streams")) else 1
        self.last_fuse = fuse
        if fuse not in (1, 2):
            raise ValueError("`fuse` must be 1 or 2.")
        if fuse == 2 and not hasattr(be, "lookahead_streams"):
            raise ValueError("this backend cannot fuse updates")
        if hasattr(be, "fuse"):
            be.fuse = fuse
        steps = (K + fuse - 1) // fuse + (self.n + RC_BLOCK - 1) // RC_BLOCK - 1
        if lookahead:
            # Look-ahead split (no collective on this path, so the steps need no host synchronisation):
            # the owner kernel of step d+1 depends only on the HEAD of step d's trail -- the two 256-row
            # blocks J, J+1 of every task -- so owner + head run on a high-priority stream while the
            # BULK of step d (blocks >= J+2) streams on a second one.  Dependencies across the streams:
            # head(d+1) reads the residual of block J+2 that bulk(d) wrote (event); bulk(d+1) needs only
            # bulk(d) (stream order) and the payloads of step d+1 (flags in the mailbox / event).
            # bulk(d) may still read its payloads while those of step d+2 arrive -- never those of
            # step d+3, because head(d+1) waits for bulk(d) before this rank's owner(d+2) can raise the
            # flags the other ranks' step d+3 depends on: four payload sets (global step & 3) suffice.
            cur = torch.cuda.current_stream(self.A.device)
            s_head, s_bulk = be.lookahead_streams()
            start = torch.cuda.Event()
            start.record(cur)
            s_head.wait_event(start)
            s_bulk.wait_event(start)
            ev_bulk = [torch.cuda.Event(), torch.cuda.Event()]
            ev_owner = [torch.cuda.Event(), torch.cuda.Event()]
            # single rank: the payloads of step d are still read by bulk(d) while owner(d+1) writes -> two buffers
            pay2 = [pay_mine, torch.empty_like(pay_mine)]
            for d in range(steps):
                with torch.cuda.stream(s_head):
                    if use_p2p:
                        be.wave_owner_p2p(d, pay_mine)
                    else:
                        be.wave_owner(d, pay2[d % 2])
                    ev_owner[d % 2].record(s_head)
                    if d > 0:
                        s_head.wait_event(ev_bulk[(d - 1) % 2])
                    if use_p2p:
                        be.wave_apply_p2p(d, 1)
                    else:
                        be.wave_apply(d, pay2[d % 2], 1)
                with torch.cuda.stream(s_bulk):
                    # bulk(d) never starts before this rank's own owner(d) has run.  Single rank: that is
                    # where the payloads come from.  Peer memory: a bulk kernel whose CTAs spin on a
                    # peer's flag may fill every SM; if it could get ahead of its own head stream, two
                    # ranks could starve each other's owner kernels (seen as a watchdog trip at n = 32768).
                    s_bulk.wait_event(ev_owner[d % 2])
                    if use_p2p:
                        be.wave_apply_p2p(d, 2)
                    else:
                        be.wave_apply(d, pay2[d % 2], 2)
                    ev_bulk[d % 2].record(s_bulk)
            done_h, done_b = torch.cuda.Event(), torch.cuda.Event()
            done_h.record(s_head)
            done_b.record(s_bulk)
            cur.wait_event(done_h)
            cur.wait_event(done_b)
            if use_p2p:
                be.wave_end_p2p(steps)
            return V_out
        if use_p2p:
            # exchange over NVLink peer memory: the owner kernel stores the coefficients into every
            # peer's mailbox, the trail kernel waits for them there -- 2 launches per step, no collective
            for d in range(steps):
                be.wave_owner_p2p(d, pay_mine)
                be.wave_apply_p2p(d)
            be.wave_end_p2p(steps)
            return V_out
        if p2p is True and multi:
            raise RuntimeError("cholupdates_b200: the peer-memory exchange is not available on this group")
        for d in range(steps):
            be.wave_owner(d, pay_mine)
            if multi:
                dist.all_gather_into_tensor(pay_all, pay_mine, group=self.group)
            be.wave_apply(d, pay_all)
        return V_out

