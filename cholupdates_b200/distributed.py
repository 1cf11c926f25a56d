"""Multi-GPU mode for batches of independent factors (BASELINE.json configs[3]).

Independent factors shard naturally: each rank (one process per GPU) owns a contiguous slice of the
batch, runs the batched kernel on it, and nothing crosses NVLink on the data path.  The only exchange
is the optional all-gather of the per-factor status vector, so that every rank can raise the same
``LinAlgError`` -- a few KB over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

The single large factor does not shard on one node's worth of memory pressure (n = 16384 fits one
GPU): ``bench.py --gpus N`` runs replicas.  The block-row-cyclic mode for n = 65536 is not built yet
(DESIGN.md section 7).
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
