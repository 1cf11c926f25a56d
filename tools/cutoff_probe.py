"""Single-factor update latency of the one-CTA kernel (path 1) vs the persistent dataflow kernel (path 3)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import _lib
lib = _lib.require_device()
dev = torch.device("cuda", 0)
status = torch.zeros(1, dtype=torch.int32, device=dev)
for n in (64, 96, 128, 192, 256, 384, 512, 768, 1024):
    L0 = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev) / np.sqrt(n), -1)
    L0.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device=dev))
    v0 = torch.randn(n, dtype=torch.float64, device=dev)
    out = []
    for path in (1, 3):
        lib.chub_set_path(path)
        ts = []
        for it in range(8):
            L, v = L0.clone(), v0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.chub_update_f64(L.data_ptr(), n, n, 0, v.data_ptr(), 0, None, 0, status.data_ptr(), None)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out.append(min(ts[2:]))
    lib.chub_set_path(0)
    print(f"n={n:5d}: one-CTA {out[0]:7.1f} us   dataflow {out[1]:7.1f} us")
