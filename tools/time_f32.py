import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import _lib
lib = _lib.require_device(); dev = torch.device("cuda", 0)
status = torch.zeros(1, dtype=torch.int32, device=dev)
for n in (4096, 16384):
    for dt, suf, es in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        L0 = torch.tril(torch.randn(n, n, dtype=dt, device=dev) / np.sqrt(n), -1)
        L0.diagonal().copy_(1 + torch.rand(n, dtype=dt, device=dev))
        v0 = torch.randn(n, dtype=dt, device=dev)
        ts = []
        for it in range(6):
            L, v = L0.clone(), v0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = getattr(lib, f"chub_update_{suf}")(L.data_ptr(), n, n, 0, v.data_ptr(), 0, None, 0, status.data_ptr(), None)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            assert rc == 0 and int(status.item()) == 0
        t = min(ts[1:]); print(f"n={n} {suf} update: {t:.3f} ms  {es * n * (n + 1) / t / 1e6:.0f} GB/s")
