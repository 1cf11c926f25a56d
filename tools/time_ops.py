"""Device-resident timings of update / downdate through the C ABI for a few sizes and layouts."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import _lib

lib = _lib.require_device()
dev = torch.device("cuda", 0)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1000, 4096, 16384]
for n in sizes:
    L0 = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev) / np.sqrt(n), -1)
    L0.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device=dev))
    u = torch.randn(n, dtype=torch.float64, device=dev); u = 0.8 * u / u.norm()
    vd0 = torch.tril(L0) @ u
    vu0 = torch.randn(n, dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    for layout in ("C", "F"):
        for op, v0 in (("update", vu0), ("downdate", vd0)):
            L = L0.clone() if layout == "C" else L0.t().contiguous().t()
            fn = getattr(lib, f"chub_{op}_f64")
            ts = []
            for it in range(6):
                Lc, v = L.clone(memory_format=torch.preserve_format), v0.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(Lc.data_ptr(), n, n, 0 if layout == "C" else 1, v.data_ptr(), 0, None, 0, status.data_ptr(), None)
                e1.record(); torch.cuda.synchronize()
                assert rc == 0 and int(status.item()) == 0, (rc, int(status.item()))
                ts.append(e0.elapsed_time(e1))
            tri = 8 * n * (n + 1) // 2
            passes = 2 if op == "update" else 3
            t = min(ts[1:])
            print(f"n={n:6d} {layout} {op:9s}: {t:8.3f} ms  ({passes * tri / t / 1e6:7.0f} GB/s on {passes} triangle passes)")
