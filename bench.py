#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

Workload (config.workload = "update_n16384_f64_inplace", BASELINE.json configs[1]): one step =
one rank-1 Seeger update ``rank_1.update(L, v, check_diag=False, overwrite_L=True,
overwrite_v=True)`` of an n = 16384 fp64 factor that is resident in HBM.  ``value`` = updates/s
over the whole job; ``ms_per_step`` = ms per update; ``roofline`` = achieved HBM GB/s of the
update against the measured copy peak (MEASURED_PEAKS.json).  The single factor does not shard
("replicas only"), so ``--gpus N`` runs N independent replicas, one per rank.

``--workload batched`` runs BASELINE.json configs[3] instead (8192 independent n = 256 factors
per GPU, sharded by rank with no collective: weak scaling).

``e2e`` is the same metric through the public API with HOST (pinned NumPy) buffers: the
host<->device copies of the lower triangle are inside the timed region.

``--impl reference`` times the reference's own CPU implementation (its Cython kernels compiled
into oracle/_ref; the C restatement oracle/ if that build is absent) on the same workload.

Contract: W untimed warm-up steps, then exactly K steps bracketed by barrier + synchronize, CUDA
events on the launching stream, max over ranks; rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_BIG = 16384
BATCH, N_SMALL = 8192, 256


def algorithmic_bytes_update(n: int, es: int = 8) -> int:
    """One read + one write of the lower triangle incl. diagonal, plus v once."""
    return 2 * es * n * (n + 1) // 2 + 2 * es * n


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------- clocks sampling
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  NVML (nvidia_ml_py) from a
    background thread every ~2 ms -- the timed region of the headline workload is only tens of
    milliseconds, too short for `nvidia-smi -lms`; nvidia-smi is the fallback when NVML is missing."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.thread = None
        self.samples = []  # (t, sm_mhz, power_w, reasons_bitmask)
        self.t_begin = self.t_end = None
        self._stop = False

    def _nvml_loop(self, nv, handle):
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    power = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                except Exception:  # noqa: BLE001
                    power = 0.0
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((time.perf_counter(), float(sm), power, int(reasons)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical GPUs: honour CUDA_VISIBLE_DEVICES if it lists indices
            idx = self.gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.gpu_index])
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.handle = nv, handle
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self.nv
            inside = [x for x in self.samples
                      if self.t_begin is not None and self.t_end is not None and self.t_begin <= x[0] <= self.t_end]
            use = inside if inside else self.samples
            if not use:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            bits = 0
            for x in use:
                bits |= x[3]
            table = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                     ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                     ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")]
            reasons = []
            for name, a, b in table:
                mask = getattr(nv, a, None) or getattr(nv, b, 0)
                if bits & int(mask):
                    reasons.append(name)
            return {"sm_mhz": statistics.median(x[1] for x in use), "sm_max_mhz": self.sm_max,
                    "samples": len(use), "in_timed_region": bool(inside), "source": "nvml",
                    "power_w_max": max(x[2] for x in use), "reasons": sorted(reasons)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "samples": len(sm), "source": "nvidia-smi",
                "power_w_max": max(power), "reasons": sorted(reasons)}


# ------------------------------------------------------------------- inputs
def make_factor_cuda(torch, n, seed, device):
    """Config-2 style factor, synthesised on the device: tril(randn / sqrt(n), -1) + diag(1 + U)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = torch.randn((n, n), dtype=torch.float64, device=device, generator=g)
    L.mul_(1.0 / np.sqrt(n))
    L = torch.tril(L, -1)
    d = 1.0 + torch.rand(n, dtype=torch.float64, device=device, generator=g)
    L.diagonal().copy_(d)
    return L


def make_factor_numpy(n, seed):
    g = np.random.default_rng(seed)
    L = np.tril(g.standard_normal((n, n)) / np.sqrt(n), -1)
    L[np.diag_indices(n)] = 1.0 + g.random(n)
    return L


# ------------------------------------------------------------------- CPU reference
def load_cpu_reference():
    """(module, kind): the reference's compiled Cython kernels if oracle/_ref is built, else the
    oracle's C restatement ("port")."""
    import build_ref

    mod = build_ref.load()
    if mod is not None:
        return mod, "reference"
    import seeger_oracle

    return seeger_oracle, "port"


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:  # noqa: BLE001
        return 1


def time_cpu_update(L_f, v_bank, reps, warmup=1):
    """Times `reps` in-place updates of the F-ordered factor L_f (the reference's fast path)."""
    mod, kind = load_cpu_reference()
    times = []
    for i in range(warmup + reps):
        v = v_bank[i % len(v_bank)].copy()
        t0 = time.perf_counter()
        mod.update(L_f, v)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "rowcyclic":
        print(json.dumps({"impl": "reference", "unavailable": "the n=65536 workload needs a 34 GB host factor and "
                          "about 2 s per update on the reference; its reference arm is --workload update"}), flush=True)
        return
    n = N_BIG if args.workload == "update" else N_SMALL
    cores = os.cpu_count()
    if args.workload == "update":
        L = np.asfortranarray(make_factor_numpy_fast(n))
        vb = [np.random.default_rng([1, i]).standard_normal(n) for i in range(4)]
        times, kind = time_cpu_update(L, vb, reps=args.steps, warmup=args.warmup)
        total = sum(times)
        value = args.steps / total
        ms = 1e3 * total / args.steps
        sample = f"{args.steps} in-place updates of one n={n} fp64 F-ordered factor (the reference's fast layout)"
        unit, metric = "updates/s", METRIC_UPDATE
        config = {"workload": "update_n16384_f64_inplace", "n": n, "layout": "F",
                  "l2": "inputs_larger_than_L2"}
    else:
        sub = 256  # bounded sample of the 8192-factor batch
        mod, kind = load_cpu_reference()
        Ls = [np.asfortranarray(make_factor_numpy(n, [0, b])) for b in range(sub)]
        vs = [np.random.default_rng([1, b]).standard_normal(n) for b in range(sub)]
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            for b in range(sub):
                mod.update(Ls[b], vs[b].copy())
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        total = sum(times)
        value = sub * args.steps / total
        ms = 1e3 * total / args.steps
        sample = f"{sub} of the {BATCH} n={n} factors per step, F-ordered"
        unit, metric = "updates/s", METRIC_BATCHED
        config = {"workload": "batched_8192x256_f64", "n": n, "batch": BATCH, "layout": "F"}
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "blas_threads": cpu_threads(),
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def make_factor_numpy_fast(n):
    """Same distribution as make_factor_numpy but with float32 draws widened (fast enough for
    n = 16384 on a few cores)."""
    g = np.random.default_rng(0)
    L = g.standard_normal((n, n), dtype=np.float32).astype(np.float64)
    L *= 1.0 / np.sqrt(n)
    L = np.tril(L, -1)
    L[np.diag_indices(n)] = 1.0 + g.random(n)
    return L


METRIC_UPDATE = "rank-1 updates/s, n=16384 fp64 in-place (ms per update = ms_per_step; HBM GB/s vs peak = roofline)"
METRIC_BATCHED = "batched rank-1 updates/s, 8192 independent n=256 fp64 factors per GPU"


# ------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import cholupdates_b200
    from cholupdates_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0 and world > 1:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.require_device()
    rank_1 = cholupdates_b200.rank_1
    peak, peak_src = measured_peaks()
    if args.path:
        lib.chub_set_path(args.path)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()  # early: the first NVML queries are slow, and the timed region is short
    stream = torch.cuda.current_stream()

    if args.workload == "update":
        n = N_BIG
        layout = args.layout
        L = make_factor_cuda(torch, n, 1234 + rank, dev)
        if layout == "F":
            L = L.t().contiguous().t()
        gv = torch.Generator(device=dev)
        gv.manual_seed(99 + rank)
        # one fresh vector per update (v is scratch on return), so the timed region is updates only
        v_bank = torch.randn((args.warmup + args.steps + 8, n), dtype=torch.float64, device=dev, generator=gv)
        units_per_step = 1
        alg_bytes = algorithmic_bytes_update(n)

        # The timed region calls the C ABI on device pointers (what rank_1.update does underneath),
        # asynchronously on the current stream; the status words are checked after the region, so
        # there is no host sync between updates.
        ws_bytes = int(lib.chub_workspace_bytes(1, n))
        ws_buf = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        status_all = torch.zeros(args.warmup + args.steps + 8, dtype=torch.int32, device=dev)
        lay = 0 if layout == "C" else 1
        call_no = [0]

        def step(i):
            rc = lib.chub_update_f64(L.data_ptr(), n, n, lay, v_bank[i].data_ptr(), 0, ws_buf.data_ptr(),
                                     ws_bytes, status_all[call_no[0]:].data_ptr(), stream.cuda_stream)
            call_no[0] += 1
            if rc != 0:
                raise RuntimeError(lib.chub_last_error().decode())

        config = {"workload": "update_n16384_f64_inplace", "n": n, "layout": layout,
                  "check_diag": False, "l2": "inputs_larger_than_L2 (1.07 GB triangle vs 126 MB L2)",
                  "replicas": world, "parallelism": "replicas only (single factor does not shard)"}
        metric = METRIC_UPDATE
    else:
        n, B = N_SMALL, BATCH
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        L = torch.randn((B, n, n), dtype=torch.float64, device=dev, generator=g).mul_(1.0 / np.sqrt(n))
        L = torch.tril(L, -1)
        L.diagonal(dim1=1, dim2=2).copy_(1.0 + torch.rand((B, n), dtype=torch.float64, device=dev, generator=g))
        v_bank = torch.randn((4, B, n), dtype=torch.float64, device=dev, generator=g)
        v = torch.empty((B, n), dtype=torch.float64, device=dev)
        units_per_step = B
        alg_bytes = B * algorithmic_bytes_update(n)

        def step(i):
            v.copy_(v_bank[i % 4])
            rank_1.update_batched(L, v, check_diag=False, overwrite_L=True, overwrite_v=True,
                                  raise_on_error=False)

        config = {"workload": "batched_8192x256_f64", "n": n, "batch_per_gpu": B, "layout": "C",
                  "l2": "inputs_larger_than_L2 (4.3 GB per GPU)", "parallelism": f"batch sharded x{world}, no collective"}
        metric = METRIC_BATCHED

    # ---- warm-up, then K timed steps (device-timed with events on the launching stream)
    for i in range(args.warmup):
        step(i)
    barrier()
    lib.chub_launch_count(1)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for i in range(args.steps):
        evs[i][0].record(stream)
        step(args.warmup + i)
        evs[i][1].record(stream)
    # (the per-step event pairs bracket one update = the library's launches: status memset,
    # pre-pass (prepare + diagonal inverses), persistent kernel)
    e1.record(stream)
    barrier()
    sampler.mark_end()
    launches = int(lib.chub_launch_count(0))
    clocks = sampler.stop()
    if args.workload == "update":
        bad = int(torch.count_nonzero(status_all).item())
        if bad:
            raise RuntimeError(f"{bad} updates reported a non-zero status")
    total_ms = e0.elapsed_time(e1)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * units_per_step * args.steps / (total_ms * 1e-3)
    ms_per_step = total_ms / args.steps
    kernel_ms = statistics.mean(step_ms)  # one update (all its launches) on the device
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/ncu_traffic.json; null if this workload has none)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(config["workload"])
        if ent:
            traffic, traffic_src = ent["traffic"], ent["source"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes": alg_bytes, "kernel_ms": kernel_ms,
                "kernel_ms_best": min(step_ms)}

    # ---- end-to-end through the public API with host (pinned) buffers
    e2e = None
    cpu_baseline = None
    if rank == 0 or world > 1:
        if args.workload == "update":
            n = N_BIG
            Lh_t = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
            Lh_t.copy_(L if args.layout == "C" else L.t())
            vh_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
            Lh, vh = Lh_t.numpy(), vh_t.numpy()
            if args.layout == "F":
                Lh = Lh.T  # the pinned buffer holds L^T row-major == L column-major
            vsrc = torch.randn((8, n), dtype=torch.float64, device=dev, generator=gv).cpu().numpy()
            reps = max(2, min(args.steps, 5))
            vh[:] = vsrc[0]
            rank_1.update(Lh, vh, check_diag=False, overwrite_L=True, overwrite_v=True)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for i in range(reps):
                vh[:] = vsrc[i % 8]
                rank_1.update(Lh, vh, check_diag=False, overwrite_L=True, overwrite_v=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            tri_bytes = 8 * n * (n + 1) // 2 + 8 * n
            e2e = {"value": world * reps / float(te.item()), "unit": "updates/s",
                   "ms_per_update": 1e3 * float(te.item()) / reps, "steps": reps,
                   "h2d_bytes_per_step": tri_bytes, "d2h_bytes_per_step": tri_bytes,
                   "api": "cholupdates_b200.rank_1.update(numpy pinned, overwrite_L=True, overwrite_v=True)"}
            if rank == 0 and world == 1 and not args.no_cpu:
                # CPU baseline: the reference's Cython path on the same factor (F order = its fast
                # path), bounded sample
                Lf = np.asfortranarray(L.cpu().numpy() if args.layout == "C" else L.t().contiguous().cpu().numpy().T)
                t_budget, times = 12.0, []
                mod, kind = load_cpu_reference()
                tstart = time.perf_counter()
                k = 0
                while True:
                    vv = vsrc[k % 8].copy()
                    t0 = time.perf_counter()
                    mod.update(Lf, vv)
                    times.append(time.perf_counter() - t0)
                    k += 1
                    if k >= 100 or time.perf_counter() - tstart > t_budget:
                        break
                times = times[1:] if len(times) > 1 else times
                cpu_baseline = {"value": 1.0 / statistics.mean(times), "unit": "updates/s",
                                "ms_per_update": 1e3 * statistics.mean(times),
                                "ms_best": 1e3 * min(times), "cores": os.cpu_count(),
                                "blas_threads": cpu_threads(), "kind": kind,
                                "sample": f"{len(times)} in-place updates of the same n={n} fp64 factor, F order (the reference's fast layout)"}
                del Lf
        else:
            # batched e2e: pinned host batch -> device -> update -> host
            n, B = N_SMALL, BATCH
            Lh = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True)
            Lh.copy_(L)
            vh = torch.empty((B, n), dtype=torch.float64, pin_memory=True)
            vh.copy_(v_bank[0])
            Ld, vd = torch.empty_like(L), torch.empty_like(v)
            reps = max(2, min(args.steps, 5))
            barrier()
            t0 = time.perf_counter()
            for i in range(reps):
                Ld.copy_(Lh, non_blocking=True)
                vd.copy_(vh, non_blocking=True)
                rank_1.update_batched(Ld, vd, check_diag=False, overwrite_L=True, overwrite_v=True,
                                      raise_on_error=False)
                Lh.copy_(Ld, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": world * B * reps / float(te.item()), "unit": "updates/s", "steps": reps,
                   "h2d_bytes_per_step": B * n * n * 8 + B * n * 8, "d2h_bytes_per_step": B * n * n * 8,
                   "api": "rank_1.update_batched on tensors copied from / to pinned host memory each step"}
            if rank == 0 and world == 1 and not args.no_cpu:
                mod, kind = load_cpu_reference()
                sub = 512
                Lc = L[:sub].cpu().numpy()
                vc = v_bank[0, :sub].cpu().numpy()
                Lf = [np.asfortranarray(Lc[b]) for b in range(sub)]
                t0 = time.perf_counter()
                for b in range(sub):
                    mod.update(Lf[b], vc[b].copy())
                dt = time.perf_counter() - t0
                cpu_baseline = {"value": sub / dt, "unit": "updates/s", "cores": os.cpu_count(),
                                "blas_threads": cpu_threads(), "kind": kind,
                                "sample": f"{sub} of the {B} n={n} factors, F order"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "step_ms": [round(x, 4) for x in step_ms[:16]],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------- row-cyclic workload
def run_rowcyclic(args):
    """BASELINE.json configs[4]: ONE n = 65536 fp64 factor distributed block-row-cyclically over the N
    GPUs of the job, 64 successive rank-1 updates per step (RowCyclicFactor.update_many: wavefront over
    (update, panel) tasks, coefficients exchanged over NVLink peer memory or NCCL).  n is fixed, so
    the scaling over N is strong."""
    import torch
    import torch.distributed as dist

    from cholupdates_b200 import _lib
    from cholupdates_b200.distributed import RowCyclicFactor, rc_owned_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    peak, peak_src = measured_peaks()
    n, K = args.n or 65536, args.k
    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows = torch.from_numpy(rc_owned_rows(n, rank, world)).to(dev)
    A = torch.empty((rows.numel(), n), dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + rank)
    cols = torch.arange(n, device=dev)
    for lo in range(0, rows.numel(), 2048):  # in chunks: no second copy of the local rows
        r = rows[lo:lo + 2048]
        blk = torch.randn((r.numel(), n), dtype=torch.float64, device=dev, generator=g).mul_(1.0 / np.sqrt(n))
        blk.masked_fill_(cols[None, :] > r[:, None], 0.0)
        blk[torch.arange(r.numel(), device=dev), r] = 1.0 + torch.rand(r.numel(), dtype=torch.float64, device=dev, generator=g)
        A[lo:lo + 2048] = blk
    del blk
    fac = RowCyclicFactor(A, n, rank, world)
    gv = torch.Generator(device=dev)
    gv.manual_seed(7)  # the same vectors on every rank
    V_bank = torch.randn((2, K, n), dtype=torch.float64, device=dev, generator=gv)
    p2p = None if args.exchange == "auto" else (args.exchange == "p2p")

    for i in range(max(args.warmup, 1)):
        fac.update_many(V_bank[i % 2], check_diag=False, p2p=p2p)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib = _lib.load()
    lib.chub_launch_count(1)
    sampler.mark_begin()
    e0.record()
    for i in range(args.steps):
        fac.update_many(V_bank[i % 2], check_diag=False, p2p=p2p)
    e1.record()
    barrier()
    sampler.mark_end()
    launches = int(lib.chub_launch_count(0))
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_update = total_ms / (args.steps * K)
    alg_bytes = algorithmic_bytes_update(n)
    per_gpu = alg_bytes / world / (ms_per_update * 1e-3) / 1e9

    # e2e: the K vectors come from pinned host memory and the K scratch vectors go back every step; the
    # 34 GB factor is the resident state of this workload (it never fits a per-step PCIe round trip)
    Vh = torch.empty((K, n), dtype=torch.float64, pin_memory=True).copy_(V_bank[0])
    Zh = torch.empty((K, n), dtype=torch.float64, pin_memory=True)
    Vd = torch.empty((K, n), dtype=torch.float64, device=dev)
    reps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for i in range(reps):
        Vd.copy_(Vh, non_blocking=True)
        Z = fac.update_many(Vd, check_diag=False, p2p=p2p)
        Zh.copy_(Z, non_blocking=True)
        torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    finite = bool(torch.isfinite(A[:4]).all().item())
    if rank == 0:
        mode = "single GPU (no exchange)" if world == 1 else (
            "peer memory (cudaIpc mailboxes)" if getattr(fac.backend, "_p2p_state", False) and p2p is not False
            else "NCCL all-gather")
        line = {
            "metric": f"successive rank-1 updates/s of one n={n} fp64 factor, block-row-cyclic over the job's GPUs ({K} per step)",
            "value": args.steps * K / (total_ms * 1e-3), "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": total_ms / args.steps,
            "ms_per_update": ms_per_update, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"rowcyclic_n{n}_k{K}_f64", "n": n, "updates_per_step": K, "exchange": mode,
                       "l2": "inputs_larger_than_L2", "parallelism": f"block-row-cyclic x{world} (256-row blocks)",
                       "result_finite": finite},
            "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s per GPU",
                         "frac": per_gpu / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes, "note": "one read + one write of the triangle per update, split over the GPUs"},
            "cpu_baseline": None,
            "e2e": {"value": reps * K / float(te.item()), "unit": "updates/s", "steps": reps,
                    "h2d_bytes_per_step": K * n * 8, "d2h_bytes_per_step": K * n * 8,
                    "api": "RowCyclicFactor.update_many(V copied from pinned host memory; scratch vectors copied back)"},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="update", choices=["update", "batched", "rowcyclic"])
    ap.add_argument("--rc-n", dest="n", type=int, default=0, help="rowcyclic: order of the factor (default 65536)")
    ap.add_argument("--rc-k", dest="k", type=int, default=64, help="rowcyclic: successive updates per step")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="rowcyclic: coefficient exchange")
    ap.add_argument("--layout", default="C", choices=["C", "F"])
    ap.add_argument("--path", type=int, default=0, help="force kernel path (1 small, 2 panel, 3 dataflow)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "rowcyclic":
        if "--steps" not in sys.argv:
            args.steps = 3
        run_rowcyclic(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
