"""End-to-end rank_1.update on pinned NumPy arrays (what bench.py's e2e leg times) for several
transfer-panel widths (env CHUB_HOST_XFER_COLS).  Usage: python tools/time_e2e.py [n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import rank_1  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda")
g.manual_seed(0)
L = torch.tril(torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n), -1)
L.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
Lh_t = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
Lh_t.copy_(L)
vh_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
Lh, vh = Lh_t.numpy(), vh_t.numpy()
vs = torch.randn((4, n), dtype=torch.float64, device="cuda", generator=g).cpu().numpy()
tri = 8 * n * (n + 1) // 2
for cols in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["256", "512", "1024", "2048", "4096"]):
    os.environ["CHUB_HOST_XFER_COLS"] = cols
    vh[:] = vs[0]
    rank_1.update(Lh, vh, check_diag=False, overwrite_L=True, overwrite_v=True)
    ts = []
    for i in range(4):
        vh[:] = vs[i]
        t0 = time.perf_counter()
        rank_1.update(Lh, vh, check_diag=False, overwrite_L=True, overwrite_v=True)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(f"xfer cols {cols:>5s}: {1e3 * t:7.2f} ms per update ({tri / t / 1e9:5.1f} GB/s each way)", flush=True)
