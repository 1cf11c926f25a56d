"""Single-GPU timing of RowCyclicFactor.update_many (world = 1): ms per update of the wavefront form
and the share of each step kernel.  Usage: python tools/time_wave.py [n] [K]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200.distributed import RowCyclicFactor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
A = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
A = torch.tril(A, -1)
A[torch.arange(n), torch.arange(n)] = 1.0 + torch.rand(n, dtype=torch.float64, device=dev, generator=g)
V = torch.randn((K, n), dtype=torch.float64, device=dev, generator=g)
fac = RowCyclicFactor(A, n)
fac.update_many(V[:2], check_diag=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e0.record()
    fac.update_many(V, check_diag=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    bytes_upd = 8 * n * (n + 1) + 16 * n
    print(f"[wave] n={n} K={K}: {ms / K:.3f} ms/update ({ms:.1f} ms total), "
          f"{bytes_upd * K / ms / 1e6:.0f} GB/s", flush=True)
assert bool(torch.isfinite(torch.tril(A)).all().item())
for la, fuse in ((False, 1), (True, 1), (True, 2)):
    e0.record()
    fac.update_many(V, check_diag=False, lookahead=la, fuse=fuse)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"[wave] n={n} K={K} lookahead={la} fuse={fuse}: {ms / K:.3f} ms/update ({ms:.1f} ms total), "
          f"{bytes_upd * K / ms / 1e6:.0f} GB/s algorithmic", flush=True)
