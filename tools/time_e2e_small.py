"""End-to-end rank_1.update / downdate on NumPy arrays at the notebook sizes (config 1), with the host-path knobs.
Usage: python tools/time_e2e_small.py [n,n,...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholupdates_b200 import rank_1  # noqa: E402

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1000, 2000, 5000]
g = np.random.default_rng(0)


def bench(label, fn, L, v0, reps=7, **kw):
    ts = []
    keep = kw.pop("_keep", False)  # pinned array: update it in place over and over (a copy would be pageable)
    for i in range(reps + 2):
        v = v0.copy()
        Lc = L.copy(order="K") if kw.get("overwrite_L") and not keep else L
        t0 = time.perf_counter()
        fn(Lc, v, check_diag=False, **kw)
        ts.append(time.perf_counter() - t0)
    ts = ts[2:]
    print(f"  {label:44s}: best {1e3 * min(ts):7.3f} ms  median {1e3 * float(np.median(ts)):7.3f} ms", flush=True)


for n in sizes:
    L0 = np.tril(g.standard_normal((n, n)) / np.sqrt(n), -1)
    L0[np.diag_indices(n)] = 1.0 + g.random(n)
    v0 = g.standard_normal(n)
    u = g.standard_normal(n); u *= 0.8 / np.linalg.norm(u)
    vd0 = L0 @ u
    for order in ("C", "F"):
        print(f"n={n} order={order}")
        Lq = np.array(L0, order=order)
        pin = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        Lp = pin.numpy() if order == "C" else pin.numpy().T
        Lp[...] = L0
        for env in ({}, {"CHUB_HOST_PATH": "streamed"}, {"CHUB_HOST_BLOCKS": "32"}, {"CHUB_HOST_BLOCKS": "128"}):
            for k in ("CHUB_HOST_PATH", "CHUB_HOST_BLOCKS"):
                os.environ.pop(k, None)
            os.environ.update(env)
            tag = " ".join(f"{k[10:]}={v}" for k, v in env.items()) or "default"
            reps = 7 if n <= 6000 else 3
            bench(f"update pageable in place [{tag}]", rank_1.update, Lq, v0, reps, overwrite_L=True, overwrite_v=True)
            if "CHUB_HOST_BLOCKS" in env:
                continue
            bench(f"update pageable copy [{tag}]", rank_1.update, Lq, v0, reps)
            bench(f"update pinned in place [{tag}]", rank_1.update, Lp, v0, reps, overwrite_L=True, overwrite_v=True, _keep=True)
            bench(f"downdate pageable in place [{tag}]", rank_1.downdate, Lq, vd0, reps, overwrite_L=True, overwrite_v=True)
        for k in ("CHUB_HOST_PATH", "CHUB_HOST_BLOCKS"):
            os.environ.pop(k, None)
        del pin, Lp
    # device-resident for scale
    Lt = torch.from_numpy(L0).cuda()
    vt = torch.from_numpy(v0).cuda()
    rank_1.update(Lt.clone(), vt.clone(), check_diag=False, overwrite_L=True, overwrite_v=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        rank_1.update(Lt, vt.clone(), check_diag=False, overwrite_L=True, overwrite_v=True)
    torch.cuda.synchronize()
    print(f"  cuda tensors, resident: {1e3 * (time.perf_counter() - t0) / 20:7.3f} ms per update")
print("cpus:", os.cpu_count())
