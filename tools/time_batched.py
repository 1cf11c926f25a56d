"""Batched kernel (8192 independent n = 256 factors, BASELINE.json configs[3]) for different
threads-per-factor settings (env CHUB_SMALL_NT = 32 | 64 | 128): ms per batch and bit-equality of
the results across settings.  Usage: python tools/time_batched.py [batch] [n] [nt,nt,...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import rank_1  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
L0 = torch.tril(torch.randn((B, n, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n), -1)
L0.diagonal(dim1=1, dim2=2).copy_(1 + torch.rand((B, n), dtype=torch.float64, device=dev, generator=g))
vu = torch.randn((B, n), dtype=torch.float64, device=dev, generator=g)
u = torch.randn((B, n), dtype=torch.float64, device=dev, generator=g)
u = 0.8 * u / u.norm(dim=1, keepdim=True)
vd = torch.einsum("bij,bj->bi", L0, u)
ref = {}
bytes_upd = B * (8 * n * (n + 1) + 16 * n)
nts = sys.argv[3].split(",") if len(sys.argv) > 3 else ["128", "64", "32"]
for nt in nts:
    os.environ["CHUB_SMALL_NT"] = nt
    for op, v0, passes in (("update", vu, 2), ("downdate", vd, 3)):
        fn = rank_1.update_batched if op == "update" else rank_1.downdate_batched
        L = L0.clone()
        ts = []
        for it in range(6):
            L.copy_(L0)
            v = v0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, status = fn(L, v, check_diag=False, overwrite_L=True, overwrite_v=True, raise_on_error=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert int(status.abs().sum()) == 0
        key = op
        if key not in ref:
            ref[key] = L.clone()
            same = True
        else:
            same = bool(torch.equal(torch.tril(ref[key]), torch.tril(L)))
        t = min(ts[1:])
        print(f"NT={nt:>3s} {op:9s}: {t:7.3f} ms/batch  {B / t / 1e3:7.2f} M/s  "
              f"({passes * bytes_upd / 2 / t / 1e6:6.0f} GB/s on {passes} triangle passes)  bit-identical to NT=128: {same}", flush=True)

# parity of the current default configuration against the staged 128-thread kernel (different
# arithmetic for the register-direct path: closed form instead of rotations -> compare by norm)
os.environ.pop("CHUB_SMALL_NT", None)
L = L0.clone()
v = vu.clone()
rank_1.update_batched(L, v, check_diag=False, overwrite_L=True, overwrite_v=True, raise_on_error=False)
torch.cuda.synchronize()
d = (torch.tril(L) - torch.tril(ref["update"])).flatten(1).norm(dim=1) / torch.tril(ref["update"]).flatten(1).norm(dim=1)
print(f"default path vs staged NT=128: max rel. Frobenius error over the batch {float(d.max()):.2e}", flush=True)
