"""End-to-end rank_1.update on pinned NumPy arrays for several lead-panel counts / panel widths of the streamed
path (env CHUB_HOST_XFER_LEAD, CHUB_HOST_XFER_COLS); the leading block is checked against the device-resident path
(CUDA tensors through the same front-end; that path is what tests/ holds against the oracle)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholupdates_b200 import rank_1  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = 2048
g = np.random.default_rng(0)
L0 = np.tril(g.standard_normal((n, n), dtype=np.float32).astype(np.float64) / np.sqrt(n), -1)
L0[np.diag_indices(n)] = 1.0 + g.random(n)
Lh_t = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
vh_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
Lh, vh = Lh_t.numpy(), vh_t.numpy()
vs = g.standard_normal((6, n))
tri = 8 * n * (n + 1) // 2
for lead, cols in ((0, 512), (1, 512), (2, 512), (4, 512), (2, 768), (4, 1024), (8, 512)):
    os.environ["CHUB_HOST_XFER_LEAD"] = str(lead)
    os.environ["CHUB_HOST_XFER_COLS"] = str(cols)
    Lh[...] = L0
    ref = torch.from_numpy(np.ascontiguousarray(L0[:m, :m])).cuda()
    ts = []
    for i in range(6):
        vh[:] = vs[i]
        t0 = time.perf_counter()
        rank_1.update(Lh, vh, check_diag=False, overwrite_L=True, overwrite_v=True)
        ts.append(time.perf_counter() - t0)
        rank_1.update(ref, torch.from_numpy(vs[i, :m].copy()).cuda(), check_diag=False, overwrite_L=True, overwrite_v=True)
    refh = ref.cpu().numpy()
    err = np.linalg.norm(np.tril(Lh[:m, :m]) - np.tril(refh)) / np.linalg.norm(np.tril(refh))
    t = min(ts[1:])
    print(f"lead {lead} cols {cols:4d}: best {1e3 * t:6.2f} ms  median {1e3 * float(np.median(ts[1:])):6.2f} ms "
          f"({tri / t / 1e9:5.1f} GB/s each way)  parity {err:.1e}", flush=True)
    assert err < 1e-12 * np.sqrt(n)
