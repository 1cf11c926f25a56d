"""Fast iteration loop for the headline kernel: n = 16384 fp64 in-place updates through the C ABI on
device pointers (what bench.py times), per-update CUDA events, plus a cross-check of the result
against the launch-per-panel path on the same input.  Usage: python tools/time_update.py [n] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cholupdates_b200 import _lib  # noqa: E402

lib = _lib.require_device()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1234)
ldpad = int(os.environ.get("CHUB_LDPAD", "0"))  # leading dimension n + ldpad (power-of-two stride effects)
ld = n + ldpad
L = torch.zeros((n, ld), dtype=torch.float64, device=dev)[:, :n]
L.copy_(torch.tril(torch.randn((n, n), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n), -1))
L.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device=dev, generator=g))
V = torch.randn((steps + 4, n), dtype=torch.float64, device=dev, generator=g)
ws_bytes = int(lib.chub_workspace_bytes(1, n))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
status = torch.zeros(steps + 4, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream

# cross-check one update against the panel path
La, Lb = L.contiguous().clone(), L.contiguous().clone()
va, vb = V[0].clone(), V[0].clone()
assert lib.chub_update_f64(La.data_ptr(), n, n, 0, va.data_ptr(), 0, ws.data_ptr(), ws_bytes, status.data_ptr(), stream) == 0
lib.chub_set_path(2)
assert lib.chub_update_f64(Lb.data_ptr(), n, n, 0, vb.data_ptr(), 0, ws.data_ptr(), ws_bytes, status[1:].data_ptr(), stream) == 0
lib.chub_set_path(0)
torch.cuda.synchronize()
err = float((torch.tril(La) - torch.tril(Lb)).norm() / torch.tril(Lb).norm())
print(f"[check] persistent vs panel path: rel. Frobenius {err:.2e}, status {status[:2].tolist()}, "
      f"upper untouched {bool((torch.triu(La, 1) == 0).all())}", flush=True)
assert os.environ.get("CHUB_NOCHECK") or (err < 1e-13 and int(status[:2].abs().sum()) == 0)
del La, Lb

Vw = V.clone()
for i in range(3):
    lib.chub_update_f64(L.data_ptr(), n, ld, 0, Vw[steps + i].data_ptr(), 0, ws.data_ptr(), ws_bytes, status.data_ptr(), stream)
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for i in range(steps):
    evs[i][0].record()
    lib.chub_update_f64(L.data_ptr(), n, ld, 0, Vw[i].data_ptr(), 0, ws.data_ptr(), ws_bytes, status[i:].data_ptr(), stream)
    evs[i][1].record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in evs]
total = evs[0][0].elapsed_time(evs[-1][1]) / steps
bytes_upd = 8 * n * (n + 1) + 16 * n
peak = 6554.2
print(f"[time] n={n}: mean {np.mean(ms):.4f} ms  median {np.median(ms):.4f}  min {min(ms):.4f}  back-to-back {total:.4f} ms/update "
      f"-> {bytes_upd / np.mean(ms) / 1e6:.0f} GB/s = {100 * bytes_upd / np.mean(ms) / 1e6 / peak:.1f} % of {peak} GB/s; "
      f"status ok: {int(status.abs().sum()) == 0}", flush=True)
