"""End-to-end rank_1.update on NumPy arrays: pinned vs pageable (bounce ring on / off), in place and copy
mode, both memory orders; results checked against the device-resident path (CUDA tensors through the same
front-end -- what tests/ holds against the oracle).  Usage: python tools/time_e2e_pageable.py [n]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholupdates_b200 import rank_1  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = np.random.default_rng(0)
L0 = np.tril(g.standard_normal((n, n), dtype=np.float32).astype(np.float64) / np.sqrt(n), -1)
L0[np.diag_indices(n)] = 1.0 + g.random(n)
vs = g.standard_normal((4, n))
tri = 8 * n * (n + 1) // 2


def bench(label, L, reps=3, **kw):
    v = vs[0].copy()
    rank_1.update(L, v, check_diag=False, **kw)
    ts = []
    for i in range(reps):
        v = vs[i % 4].copy()
        t0 = time.perf_counter()
        out = rank_1.update(L, v, check_diag=False, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"{label:58s}: {1e3 * min(ts):8.2f} ms per update ({tri / min(ts) / 1e9:5.1f} GB/s each way)", flush=True)
    return out


for order in ("C", "F"):
    pin = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
    Lp = pin.numpy() if order == "C" else pin.numpy().T
    Lp[...] = L0
    bench(f"{order} pinned, in place", Lp, overwrite_L=True, overwrite_v=True)
    Lq = np.array(L0, order=order)
    for b in ("1", "0"):
        os.environ["CHUB_HOST_BOUNCE"] = b
        bench(f"{order} pageable, in place, bounce ring {'on' if b == '1' else 'off'}", Lq, overwrite_L=True, overwrite_v=True)
    os.environ["CHUB_HOST_BOUNCE"] = "1"
    out = bench(f"{order} pageable, copy mode (overwrite_L=False)", Lq, reps=2)
    assert out is not Lq and out.flags.f_contiguous == Lq.flags.f_contiguous
    del pin, Lp
# parity of the host paths against the device-resident path
m = 4096
for order in ("C", "F"):
    junk = np.triu(np.random.default_rng(3).standard_normal((m, m)), 1)
    Lm = np.array(L0[:m, :m] + junk, order=order)
    v = vs[1, :m].copy()
    ref_t = torch.from_numpy(np.ascontiguousarray(L0[:m, :m])).cuda()
    rank_1.update(ref_t, torch.from_numpy(v.copy()).cuda(), check_diag=False, overwrite_L=True, overwrite_v=True)
    ref = ref_t.cpu().numpy()
    got = rank_1.update(Lm, v.copy(), check_diag=False)             # copy mode: fresh array, junk above the diagonal kept
    err = np.linalg.norm(np.tril(got) - np.tril(ref)) / np.linalg.norm(np.tril(ref))
    assert np.array_equal(np.triu(got, 1), junk) and np.array_equal(np.triu(Lm, 1), junk)
    rank_1.update(Lm, v, check_diag=False, overwrite_L=True, overwrite_v=True)
    err2 = np.linalg.norm(np.tril(Lm) - np.tril(ref)) / np.linalg.norm(np.tril(ref))
    assert np.array_equal(np.triu(Lm, 1), junk)
    print(f"parity n={m} {order}: copy mode {err:.2e}, in place {err2:.2e} (bar {1e-12 * np.sqrt(m):.1e}); upper triangle untouched")
    assert max(err, err2) <= 1e-12 * np.sqrt(m)
