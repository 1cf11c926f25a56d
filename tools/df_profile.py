"""Phase breakdown of the persistent dataflow kernel (cycle counters, CHUB_PROFILE=1).

Usage (on the GPU box):  CHUB_PROFILE=1 python tools/df_profile.py [n] [layout C|F]
"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CHUB_PROFILE", "1")
from cholupdates_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layout = sys.argv[2] if len(sys.argv) > 2 else "C"
lib = _lib.require_device()
lib.chub_debug_profile.restype = ctypes.c_int
lib.chub_debug_profile.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)
L = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev) / np.sqrt(n), -1)
L.diagonal().copy_(1 + torch.rand(n, dtype=torch.float64, device=dev))
if layout == "F":
    L = L.t().contiguous().t()
v = torch.randn(n, dtype=torch.float64, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
lay = 0 if layout == "C" else 1
for it in range(3):
    vv = v.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.chub_update_f64(L.data_ptr(), n, n, lay, vv.data_ptr(), 0, None, 0, status.data_ptr(), None)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, lib.chub_last_error()
print(f"n={n} layout={layout} update: {e0.elapsed_time(e1):.3f} ms, status={int(status.item())}")
ctas = 148
out = np.zeros(ctas * 16, dtype=np.int64)
lib.chub_debug_profile(None, n, out.ctypes.data, ctas)
P = out.reshape(ctas, 16)
T = (n + 31) // 32
ch = P[0]
Tc = T - T * 5 // 8 if os.environ.get("CHUB_PROFILE") == "2" else T
if not os.environ.get("CHUB_V4") and not os.environ.get("CHUB_BARRIER_CHAIN"):
    print(f"CHAIN v6 (cycles per block row over the last {Tc} of {T}): Cb total {ch[0]/Tc:.0f} | Cb: wait ring slot {ch[1]/Tc:.0f} | Cb: wait w_r {ch[3]/Tc:.0f} | "
          f"Ca: wait ring slot {ch[9]/Tc:.0f} | Ca: wait wt_r {ch[11]/Tc:.0f} | Ca: wait p_(r-1) {ch[10]/Tc:.0f} | Ca: dot + hand-over {ch[12]/Tc:.0f} | Cb: dot + publish {ch[5]/Tc:.0f} | P: poll wt {ch[4]/Tc:.0f}")
if os.environ.get("CHUB_V4"):
    print(f"CHAIN v4 (cycles per block row over the last {Tc} of {T}): C total {ch[0]/Tc:.0f} | C: wait ring slot {ch[1]/Tc:.0f} | C: M1 mat-vec {ch[4]/Tc:.0f} | "
          f"C: wait ua_r {ch[2]/Tc:.0f} | C: wait xw_r {ch[6]/Tc:.0f} | band warp 0: wait ring slot {ch[3]/Tc:.0f} | P: poll wt {ch[5]/Tc:.0f}")
print(f"CHAIN (cycles per block row over the last {Tc} of {T}): total {ch[0]/Tc:.0f} | phase A: wait for the tiles {ch[1]/Tc:.0f}, "
      f"whole phase {ch[2]/Tc:.0f} | barrier + w {ch[3]/Tc:.0f} | phase B mat-vec {ch[5]/Tc:.0f} | poll warp: wt {ch[4]/Tc:.0f}")
W = P[1:]
steps = T + 4
names = ["total", "wait tile (mbar)", "wait_read+issue load", "compute", "wait coefs", "fence+store+release", "w7 poll p", "w7 per panel total"]
for k, name in enumerate(names):
    col = W[:, k] / steps
    print(f"WORKER {name:22s}: mean {col.mean():8.0f}  min {col.min():8.0f}  max {col.max():8.0f}  cycles/step")

# ---- event trace (globaltimer, ns) of the hand-off pipeline
raw = np.zeros(256 * 16 + 8 * (n // 32 + 2), dtype=np.int64)
lib.chub_debug_profile(None, n, raw.ctypes.data, -raw.size)
tr = raw[256 * 16:][: 8 * T].reshape(8, T).astype(np.float64)
names = ["wt published (owner)", "chain sees wt", "chain publishes p", "coef warp has p", "coefs ready",
         "row warp starts step", "row warp ends step"]
t0 = tr[2][tr[2] > 0].min()
for r in (40, 41, 42, 43, 200, 201, 202, 203, 400, 401, 402):
    if r >= T: continue
    print(f"r={r:4d}: " + "  ".join(f"{names[e][:12]}={(tr[e, r] - t0) / 1e3:8.2f}us" if tr[e, r] > 0 else f"{names[e][:12]}=   --   " for e in range(7)))
# pace through the run: time per block column in windows of 32 block columns, from the chain's
# publications (event 2) and from the bottom row warp of worker 0 (event 5: iteration starts)
for ev, lab in ((2, "chain publishes p"), (3, "relay has p"), (5, "bottom row warp of worker 0 starts pan")):
    idx = np.nonzero(tr[ev] > 0)[0]
    if idx.size < 8:
        continue
    win = []
    for a in range(0, T, 32):
        sel = idx[(idx >= a) & (idx < a + 32)]
        if sel.size >= 2:
            win.append((tr[ev][sel[-1]] - tr[ev][sel[0]]) / (sel[-1] - sel[0]))
        else:
            win.append(float("nan"))
    print(f"pace [{lab}] ns per block column, windows of 32: " + " ".join(f"{w:.0f}" for w in win))
if (tr[2] > 0).sum() < 100:
    sys.exit(0)
d = np.diff(tr[2][50:T - 8])
print(f"chain period (p_r -> p_r+1): mean {d.mean():.0f} ns, median {np.median(d):.0f} ns")
for (a, b, lab) in ((0, 1, "wt published -> chain sees it"), (1, 2, "chain sees wt -> p published"), (2, 3, "p published -> coef warp has it"),
                    (3, 4, "coef warp has p -> coefs ready"), (4, 5, "coefs ready -> row warp starts step"), (5, 6, "row warp step compute")):
    m = (tr[a][50:T - 8] > 0) & (tr[b][50:T - 8] > 0)
    x = (tr[b][50:T - 8] - tr[a][50:T - 8])[m]
    print(f"{lab:40s}: mean {x.mean():8.0f} ns  median {np.median(x):8.0f}  min {x.min():8.0f}  max {x.max():8.0f}")
x = tr[5][50:T - 8] - tr[0][50:T - 8]
m = (tr[0][50:T - 8] > 0) & (tr[5][50:T - 8] > 0)
print(f"wt_r published -> row warp starts step r: mean {x[m].mean():.0f} ns")
