"""One update and one downdate per layout at n = 16384 (for ncu launch lists / captures)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import _lib
lib = _lib.require_device()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layouts = sys.argv[2] if len(sys.argv) > 2 else "CF"
ops = sys.argv[3].split(",") if len(sys.argv) > 3 else ["update", "downdate"]
dev = torch.device("cuda", 0)
L0 = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev) / np.sqrt(n), -1)
L0.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device=dev))
u = torch.randn(n, dtype=torch.float64, device=dev); u = 0.8 * u / u.norm()
vd0 = torch.tril(L0) @ u
vu0 = torch.randn(n, dtype=torch.float64, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
for layout in layouts:
    for op in ops:
        v0 = vu0 if op == "update" else vd0
        L = L0.clone() if layout == "C" else L0.t().contiguous().t()
        v = v0.clone()
        torch.cuda.synchronize()
        rc = getattr(lib, f"chub_{op}_f64")(L.data_ptr(), n, n, 0 if layout == "C" else 1, v.data_ptr(), 0, None, 0, status.data_ptr(), None)
        torch.cuda.synchronize()
        assert rc == 0 and int(status.item()) == 0
        print(layout, op, "ok")
