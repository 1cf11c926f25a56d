#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200, plus the workloads that shard.

ONE JSON line (rank 0).  Its top level is the headline workload
``update_n16384_f64_inplace`` (BASELINE.json configs[1]): one step = one rank-1 Seeger update
``rank_1.update(L, v, check_diag=False, overwrite_L=True, overwrite_v=True)`` of an n = 16384 fp64
factor resident in HBM; ``value`` = updates/s over the whole job, ``ms_per_step`` = ms per update,
``roofline`` = achieved HBM GB/s against the measured copy peak (MEASURED_PEAKS.json).  A single
n = 16384 factor does not shard ("replicas only"), so ``--gpus N`` runs N replicas of it.

``roofline.workloads`` carries the other legs of the same run, each with its own roofline and a
built-in parity check against the oracle (the driver keeps nested keys, it drops unknown top-level ones):

* ``downdate_n16384``      BASELINE.json configs[2]: in-place downdates of a resident factor (3 triangle
                           passes), the not-positive-definite case (LinAlgError, L bit-unchanged);
* ``batched_8192x256``     configs[3]: 8192 independent n = 256 factors PER GPU, sharded by rank with no
                           collective (weak scaling); status vector checked, sampled oracle comparison;
* ``rowcyclic_n65536_k64`` configs[4]: ONE n = 65536 factor block-row-cyclic over the job's GPUs, 64
                           successive updates per step (strong scaling; 34 GB fits one B200, so N = 1 is
                           a true baseline); parity: the leading 4096 x 4096 block of the updated factor
                           equals the update of the leading block with v[:4096] (computed by the oracle),
                           plus a full n = 16384 factor through the same multi-rank code path;
* ``public_api``           what a caller of ``rank_1.update`` sees: CUDA tensors (incl. the status read),
                           NumPy n = 1000 / 5000 (configs[0], the notebook workload) next to the reference.

``e2e`` is the headline metric through the public API with HOST buffers (pinned; ``e2e.pageable`` next
to it): the host<->device copies of the lower triangle are inside the timed region.

``--impl reference`` times the reference's own CPU implementation (its Cython kernels compiled into
oracle/_ref; the C restatement in oracle/ if that build is absent) on the headline workload; with
``--gpus N`` rank 0 runs N replica processes on the host's cores (the whole host against the whole job).

``--workload update|downdate|batched|rowcyclic`` runs a single leg as the top-level line (development).

Contract: W untimed warm-up steps, then exactly K steps bracketed by barrier + synchronize, CUDA
events on the launching stream, max over ranks; rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_BIG = 16384
BATCH, N_SMALL = 8192, 256
RC_N, RC_K = 65536, 64
LEAD = 4096  # order of the leading block the oracle re-computes for the big factors

METRIC_UPDATE = "rank-1 updates/s, n=16384 fp64 in-place (ms per update = ms_per_step; HBM GB/s vs peak = roofline)"
METRIC_BATCHED = "batched rank-1 updates/s, 8192 independent n=256 fp64 factors per GPU"


def algorithmic_bytes_update(n: int, es: int = 8) -> int:
    """One read + one write of the lower triangle incl. diagonal, plus v once."""
    return 2 * es * n * (n + 1) // 2 + 2 * es * n


def algorithmic_bytes_downdate(n: int, es: int = 8) -> int:
    """Three triangle passes (trsv read; sweep read + write), SURVEY 8d."""
    return 3 * es * n * (n + 1) // 2 + 2 * es * n


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


def make_factor_numpy_fast(n, seed=0):
    """Same distribution as make_factor_numpy but with float32 draws widened (fast enough for
    n = 16384 on a few cores)."""
    g = np.random.default_rng(seed)
    L = g.standard_normal((n, n), dtype=np.float32).astype(np.float64)
    L *= 1.0 / np.sqrt(n)
    L = np.tril(L, -1)
    L[np.diag_indices(n)] = 1.0 + g.random(n)
    return L


def config1_inputs(n):
    """BASELINE.json configs[0] / notebooks/benchmark_rank_1.ipynb:30-33,49: the
    reference's ``random_spd_matrix(n, fast=True, rng=default_rng(0))`` (utils.py:78-90, restated:
    the reference package itself does not travel to the GPU box) + ``scipy.linalg.cholesky(lower=True)``
    and ``v ~ N(0, 10^2)``.  tests/test_oracle.py pins the restatement against the live generator."""
    import scipy.linalg

    g = np.random.default_rng(0)
    A = g.standard_normal((n, n))
    M = A @ A.T
    M += np.eye(n)
    D = np.sqrt(np.diag(M))
    M = D[:, None] * M * D[None, :]
    M = 0.5 * (M + M.T)
    L = scipy.linalg.cholesky(M, lower=True)
    v = np.random.default_rng(1).normal(scale=10.0, size=n)
    return L, v


# ------------------------------------------------------------------- CPU reference / oracle
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


def oracle_apply(L0, ops):
    """Apply ``ops`` = [("update" | "downdate", v), ...] to a copy of L0 with the CPU oracle (F order,
    its fast layout).  Only used as the CHECKER of the legs below."""
    mod, _ = load_cpu_reference()
    L = np.asfortranarray(np.tril(L0))
    for op, v in ops:
        getattr(mod, op)(L, np.ascontiguousarray(v, dtype=L.dtype).copy())
    return L


def parity_report(L_gpu, L_ref, n_tol=None):
    """Relative Frobenius error over the lower triangle against the oracle; bar 1e-12 * sqrt(n)."""
    n = L_ref.shape[0]
    a, b = np.tril(np.asarray(L_gpu, dtype=np.float64)), np.tril(np.asarray(L_ref, dtype=np.float64))
    err = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    tol = 1e-12 * np.sqrt(n_tol or n)
    return {"rel_fro_vs_oracle": err, "tol": tol, "ok": bool(err <= tol)}


# ------------------------------------------------------------------- reference arm
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


def reference_worker(args):
    """One replica of the reference arm: prints {"total_s": ..., "kind": ...} for `steps` updates."""
    n = N_BIG
    L = np.asfortranarray(make_factor_numpy_fast(n, seed=args.seed))
    vb = [np.random.default_rng([1, args.seed, i]).standard_normal(n) for i in range(4)]
    # start barrier over a file so that the replicas overlap
    if args.sync_file:
        open(args.sync_file + f".ready{args.seed}", "w").close()
        t_dead = time.time() + 300
        while not os.path.exists(args.sync_file + ".go") and time.time() < t_dead:
            time.sleep(0.01)
    times, kind = time_cpu_update(L, vb, reps=args.steps, warmup=args.warmup)
    print(json.dumps({"total_s": sum(times), "kind": kind, "threads": cpu_threads()}), flush=True)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = N_BIG
    cores = os.cpu_count() or 1
    replicas = max(1, args.gpus)
    if replicas == 1:
        L = np.asfortranarray(make_factor_numpy_fast(n))
        vb = [np.random.default_rng([1, i]).standard_normal(n) for i in range(4)]
        times, kind = time_cpu_update(L, vb, reps=args.steps, warmup=args.warmup)
        total, threads = sum(times), cpu_threads()
        sample = f"{args.steps} in-place updates of one n={n} fp64 F-ordered factor (the reference's fast layout)"
    else:
        # the whole host against the whole job: N replica processes, the BLAS threads split between them
        import tempfile

        per = max(1, cores // replicas)
        sync = os.path.join(tempfile.mkdtemp(prefix="chub_ref_"), "sync")
        env = dict(os.environ, OPENBLAS_NUM_THREADS=str(per), OMP_NUM_THREADS=str(per), MKL_NUM_THREADS=str(per))
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--impl", "reference-worker",
                                   "--steps", str(args.steps), "--warmup", str(args.warmup), "--seed", str(i),
                                   "--sync-file", sync], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                  text=True, env=env) for i in range(replicas)]
        t_dead = time.time() + 300
        while not all(os.path.exists(sync + f".ready{i}") for i in range(replicas)) and time.time() < t_dead:
            time.sleep(0.05)
        open(sync + ".go", "w").close()
        outs = [json.loads(p.communicate(timeout=1800)[0].strip().splitlines()[-1]) for p in procs]
        total = max(o["total_s"] for o in outs)
        kind, threads = outs[0]["kind"], outs[0]["threads"]
        sample = (f"{replicas} replica processes x {args.steps} in-place updates of an n={n} fp64 F-ordered factor, "
                  f"{per} BLAS threads each, running concurrently on the host's {cores} cores")
    value = replicas * args.steps / total
    ms = 1e3 * total / args.steps
    line = {
        "impl": "reference", "metric": METRIC_UPDATE, "value": value, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "update_n16384_f64_inplace", "n": n, "layout": "F", "replicas": replicas,
                   "l2": "inputs_larger_than_L2"},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": cores, "blas_threads": threads,
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------- GPU arm: context
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import cholupdates_b200
        from cholupdates_b200 import _lib

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.rank == 0 and self.world > 1:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={self.world}", file=sys.stderr)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
        self.lib = _lib.require_device()
        self.rank_1 = cholupdates_b200.rank_1
        self.peak, self.peak_src = measured_peaks()
        if args.path:
            self.lib.chub_set_path(args.path)
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def traffic(self, workload):
        """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        (profiles/ncu_traffic.json; None if this workload has none)."""
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ent = json.load(f).get(workload)
            if ent:
                return ent["traffic"], ent["source"]
        except (OSError, ValueError, KeyError):
            pass
        return None, None


def timed_steps(ctx, step, warmup, steps, sampler=None):
    """W warm-ups, then K steps: one event pair per step + one around all of them; barrier +
    synchronize on both sides; returns (total_ms max over ranks, per-step ms of this rank)."""
    torch = ctx.torch
    for i in range(warmup):
        step(i)
    ctx.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    if sampler:
        sampler.mark_begin()
    e0.record(ctx.stream)
    for i in range(steps):
        evs[i][0].record(ctx.stream)
        step(warmup + i)
        evs[i][1].record(ctx.stream)
    e1.record(ctx.stream)
    ctx.barrier()
    if sampler:
        sampler.mark_end()
    total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    return total_ms, [a.elapsed_time(b) for a, b in evs]


# ------------------------------------------------------------------- leg: headline update
def leg_update(ctx, args, sampler):
    torch, lib, rank_1 = ctx.torch, ctx.lib, ctx.rank_1
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n, layout = N_BIG, args.layout
    L = make_factor_cuda(torch, n, 1234 + rank, dev)
    if layout == "F":
        L = L.t().contiguous().t()
    gv = torch.Generator(device=dev)
    gv.manual_seed(99 + rank)
    # one fresh vector per update (v is scratch on return), so the timed region is updates only
    nv = args.warmup + args.steps
    v_bank = torch.randn((nv + 8, n), dtype=torch.float64, device=dev, generator=gv)
    v_lead = v_bank[:nv, :LEAD].cpu().numpy()  # what the oracle applies to the leading block
    L_lead0 = L[:LEAD, :LEAD].cpu().numpy()
    alg_bytes = algorithmic_bytes_update(n)

    # The timed region calls the C ABI on device pointers (what rank_1.update does underneath),
    # asynchronously on the current stream; the status words are checked after the region, so
    # there is no host sync between updates (public_api times the call a user makes, sync included).
    ws_bytes = int(lib.chub_workspace_bytes(1, n))
    ws_buf = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    status_all = torch.zeros(nv + 8, dtype=torch.int32, device=dev)
    lay = 0 if layout == "C" else 1
    call_no = [0]

    def step(i):
        rc = lib.chub_update_f64(L.data_ptr(), n, n, lay, v_bank[i].data_ptr(), 0, ws_buf.data_ptr(),
                                 ws_bytes, status_all[call_no[0]:].data_ptr(), ctx.stream.cuda_stream)
        call_no[0] += 1
        if rc != 0:
            raise RuntimeError(lib.chub_last_error().decode())

    for i in range(args.warmup):
        step(i)
    ctx.barrier()
    lib.chub_launch_count(1)
    total_ms, step_ms = timed_steps(ctx, lambda i: step(args.warmup + i), 0, args.steps, sampler)
    # (warm-up ran above so that the launch counter covers exactly the timed region)
    launches = int(lib.chub_launch_count(0))
    clocks = sampler.stop() if sampler else None
    bad = int(torch.count_nonzero(status_all).item())
    if bad:
        raise RuntimeError(f"{bad} updates reported a non-zero status")
    value = world * args.steps / (total_ms * 1e-3)
    kernel_ms = statistics.mean(step_ms)  # one update (all its launches) on the device
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ctx.traffic("update_n16384_f64_inplace")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": ctx.peak, "unit": "GB/s",
                "frac": achieved / ctx.peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": ctx.peak_src, "algorithmic_bytes": alg_bytes, "kernel_ms": kernel_ms,
                "kernel_ms_best": min(step_ms)}

    # parity: the leading block of the updated factor == the update of the leading block (oracle)
    parity = None
    if rank == 0:
        L_ref = oracle_apply(L_lead0, [("update", v_lead[i]) for i in range(nv)])
        parity = parity_report(L[:LEAD, :LEAD].cpu().numpy(), L_ref)
        parity["what"] = (f"leading {LEAD} x {LEAD} block after {nv} successive updates vs the oracle applied to the "
                          "leading block (the update of a leading block depends on that block and v[:m] only)")
    roofline["parity"] = parity

    # ---- end-to-end through the public API with host buffers (copies inside the timed region)
    e2e, cpu_baseline = None, None
    Lh_t = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
    Lh_t.copy_(L if layout == "C" else L.t())
    vh_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
    Lh, vh = Lh_t.numpy(), vh_t.numpy()
    if layout == "F":
        Lh = Lh.T  # the pinned buffer holds L^T row-major == L column-major
    vsrc = torch.randn((8, n), dtype=torch.float64, device=dev, generator=gv).cpu().numpy()
    reps = max(2, min(args.steps, 5))

    def host_loop(Lx, vx, reps):
        vx[:] = vsrc[0]
        rank_1.update(Lx, vx, check_diag=False, overwrite_L=True, overwrite_v=True)  # warm-up
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(reps):
            vx[:] = vsrc[i % 8]
            rank_1.update(Lx, vx, check_diag=False, overwrite_L=True, overwrite_v=True)
        torch.cuda.synchronize()
        return ctx.max_over_ranks(time.perf_counter() - t0)

    dt = host_loop(Lh, vh, reps)
    tri_bytes = 8 * n * (n + 1) // 2 + 8 * n
    e2e = {"value": world * reps / dt, "unit": "updates/s", "ms_per_update": 1e3 * dt / reps, "steps": reps,
           "h2d_bytes_per_step": tri_bytes, "d2h_bytes_per_step": tri_bytes,
           "api": "cholupdates_b200.rank_1.update(numpy pinned, overwrite_L=True, overwrite_v=True)"}
    if not args.quick:
        # the same call on an ordinary (pageable) ndarray -- what a drop-in caller passes
        Lp = np.array(Lh, order="K", copy=True)
        vp = np.empty(n)
        dtp = host_loop(Lp, vp, max(2, reps // 2))
        e2e["pageable"] = {"value": world * max(2, reps // 2) / dtp, "ms_per_update": 1e3 * dtp / max(2, reps // 2),
                           "api": "the same call on a pageable numpy.ndarray"}
        del Lp
        # successive updates of a factor that stays on the device (resident handle): only v travels
        try:
            e2e["resident"] = e2e_resident(ctx, Lh, vsrc, reps)
        except Exception as exc:  # noqa: BLE001
            e2e["resident"] = {"error": repr(exc)[:200]}
    if rank == 0 and world == 1 and not args.no_cpu:
        # CPU baseline: the reference's Cython path on the same factor (F order = its fast path)
        Lf = np.asfortranarray(L.cpu().numpy() if layout == "C" else L.t().contiguous().cpu().numpy().T)
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
                        "ms_per_update": 1e3 * statistics.mean(times), "ms_best": 1e3 * min(times),
                        "cores": os.cpu_count(), "blas_threads": cpu_threads(), "kind": kind,
                        "sample": f"{len(times)} in-place updates of the same n={n} fp64 factor, F order (the reference's fast layout)"}
        del Lf
    del Lh_t, L
    torch.cuda.empty_cache()
    config = {"workload": "update_n16384_f64_inplace", "n": n, "layout": layout, "check_diag": False,
              "l2": "inputs_larger_than_L2 (1.07 GB triangle vs 126 MB L2)", "replicas": world,
              "parallelism": "replicas only (single factor does not shard)"}
    return {"metric": METRIC_UPDATE, "value": value, "unit": "updates/s", "ms_per_step": total_ms / args.steps,
            "scaling": "weak", "config": config, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "step_ms": [round(x, 4) for x in step_ms[:16]]}


def e2e_resident(ctx, Lh, vsrc, reps):
    """Successive updates of ONE host factor through a resident-factor handle: the factor is uploaded
    once, every update ships only v (host -> device) and the scratch vector back, and the factor is
    downloaded once at the end.  Reported per update, upload / download amortised over `reps`."""
    from cholupdates_b200 import resident

    torch = ctx.torch
    n = Lh.shape[0]
    ctx.barrier()
    t0 = time.perf_counter()
    with resident.ResidentFactor(Lh) as fac:
        for i in range(reps):
            fac.update(vsrc[i % 8].copy(), check_diag=False)
        fac.download(out=Lh)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    # steady state: the updates alone (factor already resident)
    with resident.ResidentFactor(Lh) as fac:
        fac.update(vsrc[0].copy(), check_diag=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(4 * reps):
            fac.update(vsrc[i % 8].copy(), check_diag=False)
        torch.cuda.synchronize()
        dts = ctx.max_over_ranks(time.perf_counter() - t0)
    return {"ms_per_update_incl_upload_download": 1e3 * dt / reps, "updates": reps,
            "ms_per_update_steady_state": 1e3 * dts / (4 * reps),
            "h2d_bytes_per_update_steady_state": 8 * n, "d2h_bytes_per_update_steady_state": 8 * n + 4,
            "api": "cholupdates_b200.resident.ResidentFactor(L).update(v) ... .download()"}


# ------------------------------------------------------------------- leg: downdate n = 16384
def leg_downdate(ctx, args):
    torch, lib, rank_1 = ctx.torch, ctx.lib, ctx.rank_1
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n, layout = N_BIG, args.layout
    steps, warmup = max(3, min(args.steps, 10)), 3
    L = make_factor_cuda(torch, n, 4321 + rank, dev)
    if layout == "F":
        L = L.t().contiguous().t()
    g = torch.Generator(device=dev)
    g.manual_seed(17 + rank)
    U = torch.randn((warmup + steps + 2, n), dtype=torch.float64, device=dev, generator=g)
    U = U / U.norm(dim=1, keepdim=True)
    ws_bytes = int(lib.chub_workspace_bytes(1, n))
    ws_buf = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    status_all = torch.zeros(warmup + steps + 4, dtype=torch.int32, device=dev)
    lay = 0 if layout == "C" else 1
    L_lead0 = L[:LEAD, :LEAD].cpu().numpy()
    applied = []
    vbuf = torch.empty(n, dtype=torch.float64, device=dev)
    evs = []

    def downdate(i, timed):
        # v = L (alpha u): p = alpha u, rho^2 = 1 - alpha^2; alpha^2 = 0.25 keeps
        # the factor well conditioned over the successive downdates
        torch.mv(L, 0.5 * U[i], out=vbuf)  # (the strict upper triangle of L is zero here)
        applied.append(vbuf[:LEAD].cpu().numpy().copy())
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
        rc = lib.chub_downdate_f64(L.data_ptr(), n, n, lay, vbuf.data_ptr(), 0, ws_buf.data_ptr(), ws_bytes,
                                   status_all[i:].data_ptr(), ctx.stream.cuda_stream)
        if timed:
            e1.record(ctx.stream)
            evs.append((e0, e1))
        if rc != 0:
            raise RuntimeError(lib.chub_last_error().decode())

    for i in range(warmup):
        downdate(i, False)
    ctx.barrier()
    for i in range(steps):
        downdate(warmup + i, True)
    ctx.barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    ms = ctx.max_over_ranks(statistics.mean(step_ms))
    if int(torch.count_nonzero(status_all).item()):
        raise RuntimeError("a downdate reported a non-zero status")
    b3, b2 = algorithmic_bytes_downdate(n), algorithmic_bytes_update(n)
    out = {"workload": "downdate_n16384_f64_inplace", "n": n, "layout": layout, "steps": steps, "warmup": warmup,
           "ms_per_downdate": ms, "ms_best": min(step_ms), "value": world / (ms * 1e-3), "unit": "downdates/s",
           "roofline": {"bound": "hbm", "achieved": b3 / (ms * 1e-3) / 1e9, "peak": ctx.peak, "unit": "GB/s",
                        "frac": b3 / (ms * 1e-3) / 1e9 / ctx.peak, "algorithmic_bytes": b3,
                        "frac_of_2_pass_bytes": b2 / (ms * 1e-3) / 1e9 / ctx.peak,
                        "note": "3 triangle passes: forward substitution (read), sweep (read + write)"}}
    if rank == 0:
        L_ref = oracle_apply(L_lead0, [("downdate", v) for v in applied])
        out["parity"] = parity_report(L[:LEAD, :LEAD].cpu().numpy(), L_ref)
        out["parity"]["what"] = (f"leading {LEAD} x {LEAD} block after {len(applied)} successive downdates vs the oracle "
                                 "downdating the leading block with v[:m] (Cholesky factor of the leading principal minor)")
    # the error path of configs[2]: ||L^{-1} v|| = 1.2 -> LinAlgError, L bit-unchanged, v <- p
    torch.mv(L, 1.2 * U[warmup + steps], out=vbuf)
    before = L.clone()
    raised = False
    try:
        rank_1.downdate(L, vbuf, check_diag=True, overwrite_L=True, overwrite_v=True)
    except np.linalg.LinAlgError:
        raised = True
    p_err = float((vbuf - 1.2 * U[warmup + steps]).norm() / 1.2)
    out["not_pd_case"] = {"raised_LinAlgError": raised, "L_bit_unchanged": bool(torch.equal(L, before)),
                          "v_holds_p_rel_err": p_err, "ok": bool(raised and torch.equal(L, before) and p_err < 1e-9)}
    del before, L
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------- leg: other sizes / the other layout
def leg_update_sized(ctx, args, n, layout, steps=10, warmup=3):
    """In-place updates of one resident factor of order n in `layout` through the C ABI (device pointers): the
    column-major layout the reference recommends (_seeger.py:65-68) and sizes beyond one persistent launch
    (blocked driver).  Parity: leading block vs the oracle, as in the headline leg."""
    torch, lib = ctx.torch, ctx.lib
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    L = make_factor_cuda(torch, n, 777 + rank + n, dev)
    if layout == "F":
        L = L.t().contiguous().t()
    g = torch.Generator(device=dev)
    g.manual_seed(29 + rank)
    V = torch.randn((warmup + steps, n), dtype=torch.float64, device=dev, generator=g)
    V0 = V[:, :LEAD].cpu().numpy().copy()
    ws_bytes = int(lib.chub_workspace_bytes(1, n))
    ws_buf = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    status_all = torch.zeros(warmup + steps, dtype=torch.int32, device=dev)
    lay = 0 if layout == "C" else 1
    L_lead0 = L[:LEAD, :LEAD].cpu().numpy()

    def step(i):
        rc = lib.chub_update_f64(L.data_ptr(), n, n, lay, V[i].data_ptr(), 0, ws_buf.data_ptr(), ws_bytes,
                                 status_all[i:].data_ptr(), ctx.stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(lib.chub_last_error().decode())

    _, step_ms = timed_steps(ctx, step, warmup, steps)
    ms = ctx.max_over_ranks(statistics.mean(step_ms))
    if int(torch.count_nonzero(status_all).item()):
        raise RuntimeError("an update reported a non-zero status")
    b2 = algorithmic_bytes_update(n)
    out = {"workload": f"update_n{n}_f64_inplace_{layout}", "n": n, "layout": layout, "steps": steps, "warmup": warmup,
           "ms_per_update": ms, "ms_best": min(step_ms), "value": world / (ms * 1e-3), "unit": "updates/s",
           "roofline": {"bound": "hbm", "achieved": b2 / (ms * 1e-3) / 1e9, "peak": ctx.peak, "unit": "GB/s",
                        "frac": b2 / (ms * 1e-3) / 1e9 / ctx.peak, "algorithmic_bytes": b2}}
    tr, src = ctx.traffic(out["workload"])
    if tr:
        out["roofline"]["traffic"], out["roofline"]["traffic_source"] = tr, src
    if rank == 0:
        L_ref = oracle_apply(L_lead0, [("update", v) for v in V0])
        out["parity"] = parity_report(L[:LEAD, :LEAD].cpu().numpy(), L_ref, n_tol=n)
        out["parity"]["what"] = (f"leading {LEAD} x {LEAD} block after {warmup + steps} successive updates vs the oracle "
                                 "applied to the leading block")
    del L, V, ws_buf
    torch.cuda.empty_cache()
    return out


def leg_layouts_and_sizes(ctx, args):
    """Column-major n = 16384 (update and downdate) and n = 32768 in both layouts (two diagonal blocks + one
    rectangle each: the blocked driver)."""
    import copy

    out = {}
    fargs = copy.copy(args)
    fargs.layout = "F"
    out["update_n16384_F"] = leg_update_sized(ctx, args, N_BIG, "F", steps=20)
    dd = leg_downdate(ctx, fargs)
    dd.pop("not_pd_case", None)
    out["downdate_n16384_F"] = dd
    out["update_n32768_C"] = leg_update_sized(ctx, args, 32768, "C", steps=5)
    out["update_n32768_F"] = leg_update_sized(ctx, args, 32768, "F", steps=5)
    return out


# ------------------------------------------------------------------- leg: batched 8192 x 256 per GPU
def leg_batched(ctx, args):
    torch, rank_1 = ctx.torch, ctx.rank_1
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n, B = N_SMALL, BATCH
    steps, warmup = max(3, min(args.steps, 20)), 3
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    L = torch.randn((B, n, n), dtype=torch.float64, device=dev, generator=g).mul_(1.0 / np.sqrt(n))
    L = torch.tril(L, -1)
    L.diagonal(dim1=1, dim2=2).copy_(1.0 + torch.rand((B, n), dtype=torch.float64, device=dev, generator=g))
    v_bank = torch.randn((4, B, n), dtype=torch.float64, device=dev, generator=g)
    v = torch.empty((B, n), dtype=torch.float64, device=dev)
    sample = sorted(set(list(range(8)) + list(range(B // 2, B // 2 + 8)) + list(range(B - 8, B))))
    L0 = L[sample].cpu().numpy()
    status_sum = torch.zeros((), dtype=torch.int64, device=dev)

    def step(i):
        v.copy_(v_bank[i % 4])
        _, st = rank_1.update_batched(L, v, check_diag=False, overwrite_L=True, overwrite_v=True,
                                      raise_on_error=False)
        status_sum.add_(st.ne(0).sum())  # device-side: no host sync inside the timed region

    total_ms, step_ms = timed_steps(ctx, step, warmup, steps)
    failed = int(status_sum.item())
    ms = total_ms / steps
    alg = B * algorithmic_bytes_update(n)
    achieved = alg / (statistics.mean(step_ms) * 1e-3) / 1e9
    traffic, traffic_src = ctx.traffic("batched_8192x256_f64")
    out = {"workload": "batched_8192x256_f64", "n": n, "batch_per_gpu": B, "steps": steps, "warmup": warmup,
           "ms_per_batch": ms, "value": world * B / (ms * 1e-3), "unit": "updates/s", "scaling": "weak",
           "parallelism": f"batch sharded x{world}, no collective", "status_nonzero": failed,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": ctx.peak, "unit": "GB/s per GPU",
                        "frac": achieved / ctx.peak, "algorithmic_bytes": alg, "traffic": traffic,
                        "traffic_source": traffic_src}}
    # sampled oracle comparison: the same warmup + steps vectors applied to 24 of the 8192 factors
    vb = v_bank[:, sample].cpu().numpy()
    worst = 0.0
    for k in range(len(sample)):
        ref = oracle_apply(L0[k], [("update", vb[i % 4, k]) for i in range(warmup + steps)])
        worst = max(worst, parity_report(L[sample[k]].cpu().numpy(), ref)["rel_fro_vs_oracle"])
    tol = 1e-12 * np.sqrt(n)
    ok = ctx.all_ok(worst <= tol and failed == 0)
    out["parity"] = {"rel_fro_vs_oracle_worst": worst, "tol": tol, "ok": ok, "factors_checked": len(sample),
                     "what": f"{len(sample)} sampled factors per rank after {warmup + steps} successive updates vs the oracle; "
                             "status vector all zero on every rank"}
    del L, v_bank
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------- leg: row-cyclic n = 65536, K = 64
def rc_make_local(ctx, n, seed0):
    """This rank's rows of the config-5 factor: 256-row block r is drawn from a generator seeded by r,
    so that any rank can regenerate any block (the parity check regenerates the leading ones)."""
    from cholupdates_b200.distributed import RC_BLOCK

    torch, dev = ctx.torch, ctx.dev
    nblk = (n + RC_BLOCK - 1) // RC_BLOCK
    mine = list(range(ctx.rank, nblk, ctx.world))
    rows = sum(min(RC_BLOCK, n - b * RC_BLOCK) for b in mine)
    A = torch.empty((rows, n), dtype=torch.float64, device=dev)
    lo = 0
    for b in mine:
        blk = rc_block(ctx, n, b, seed0)
        A[lo:lo + blk.shape[0]] = blk
        lo += blk.shape[0]
    return A


def rc_block(ctx, n, b, seed0):
    from cholupdates_b200.distributed import RC_BLOCK

    torch, dev = ctx.torch, ctx.dev
    g = torch.Generator(device=dev)
    g.manual_seed(seed0 + b)
    r = torch.arange(b * RC_BLOCK, min(n, (b + 1) * RC_BLOCK), device=dev)
    blk = torch.randn((r.numel(), n), dtype=torch.float64, device=dev, generator=g).mul_(1.0 / np.sqrt(n))
    cols = torch.arange(n, device=dev)
    blk.masked_fill_(cols[None, :] > r[:, None], 0.0)
    blk[torch.arange(r.numel(), device=dev), r] = 1.0 + torch.rand(r.numel(), dtype=torch.float64, device=dev, generator=g)
    return blk


def rc_gather_leading(ctx, A, n, m):
    """The leading m x m block of the distributed factor on every rank (sum of the ranks' owned rows)."""
    from cholupdates_b200.distributed import rc_owned_rows

    torch = ctx.torch
    rows = rc_owned_rows(n, ctx.rank, ctx.world)
    out = torch.zeros((m, m), dtype=torch.float64, device=ctx.dev)
    sel = np.nonzero(rows < m)[0]
    if sel.size:
        idx = torch.from_numpy(rows[sel]).to(ctx.dev)
        out[idx] = A[torch.from_numpy(sel).to(ctx.dev), :m]
    if ctx.world > 1:
        ctx.dist.all_reduce(out)
    return out


def leg_rowcyclic(ctx, args, n=None, K=None):
    from cholupdates_b200.distributed import RC_BLOCK, RowCyclicFactor

    torch, lib = ctx.torch, ctx.lib
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n, K = n or args.n or RC_N, K or args.k
    steps, warmup = max(1, min(args.steps, 2 if n >= 32768 else 4)), 1
    p2p = None if args.exchange == "auto" else (args.exchange == "p2p")
    seed0 = 50000
    A = rc_make_local(ctx, n, seed0)
    fac = RowCyclicFactor(A, n, rank, world)
    gv = torch.Generator(device=dev)
    gv.manual_seed(7)  # the same vectors on every rank
    V_bank = torch.randn((2, K, n), dtype=torch.float64, device=dev, generator=gv)
    m = min(LEAD, n)
    lead0 = torch.cat([rc_block(ctx, n, b, seed0)[:, :m] for b in range((m + RC_BLOCK - 1) // RC_BLOCK)])[:m].cpu().numpy()

    fuse = args.fuse if args.fuse else None

    def step(i):
        fac.update_many(V_bank[i % 2], check_diag=False, p2p=p2p, fuse=fuse)

    for i in range(warmup):
        step(i)
    ctx.barrier()
    lib.chub_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for i in range(steps):
        step(warmup + i)
    e1.record(ctx.stream)
    ctx.barrier()
    launches = int(lib.chub_launch_count(0))
    total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_per_update = total_ms / (steps * K)
    alg = algorithmic_bytes_update(n)
    used_fuse = int(getattr(fac, "last_fuse", 1) or 1)
    # fused passes: two updates share one read + write of a panel, so the bytes actually MOVED per update
    # are alg / fuse; the roofline fraction is quoted on moved bytes (never above 1), the algorithmic rate
    # is reported next to it
    per_gpu_alg = alg / world / (ms_per_update * 1e-3) / 1e9
    per_gpu = per_gpu_alg / used_fuse
    mode = "single GPU (no exchange)" if world == 1 else (
        "peer memory (cudaIpc mailboxes, flags in the kernels)" if getattr(fac.backend, "_p2p_state", False) and p2p is not False
        else "NCCL all-gather per step")
    out = {"workload": f"rowcyclic_n{n}_k{K}_f64", "n": n, "updates_per_step": K, "steps": steps, "warmup": warmup,
           "ms_per_update": ms_per_update, "ms_per_step": total_ms / steps, "value": steps * K / (total_ms * 1e-3),
           "unit": "updates/s", "scaling": "strong", "exchange": mode, "fuse": used_fuse,
           "parallelism": f"block-row-cyclic x{world} (256-row blocks)", "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": ctx.peak, "unit": "GB/s per GPU (bytes moved)",
                        "frac": per_gpu / ctx.peak, "algorithmic_bytes": alg, "moved_bytes_per_update": alg // used_fuse,
                        "algorithmic_gbs_per_gpu": per_gpu_alg, "updates_per_pass": used_fuse,
                        "note": "one read + one write of the triangle per PASS, split over the GPUs; with "
                                "updates_per_pass = 2 two successive updates share a pass"}}
    # parity 1: leading block of the big factor vs the oracle (all warmup + steps calls, K vectors each)
    lead = rc_gather_leading(ctx, A, n, m).cpu().numpy()
    if rank == 0:
        Vh = V_bank[:, :, :m].cpu().numpy()
        ops = [("update", Vh[i % 2, u]) for i in range(warmup + steps) for u in range(K)]
        par = parity_report(lead, oracle_apply(lead0, ops))
        par["what"] = (f"leading {m} x {m} block (gathered from the ranks that own its rows) after {len(ops)} successive "
                       "updates vs the oracle applied to the leading block")
        out["parity"] = par
    # e2e: the K vectors come from pinned host memory and the K scratch vectors go back every step;
    # the factor is the resident state of this workload (34 GB never fits a per-step PCIe round trip)
    Vh_p = torch.empty((K, n), dtype=torch.float64, pin_memory=True).copy_(V_bank[0])
    Zh = torch.empty((K, n), dtype=torch.float64, pin_memory=True)
    Vd = torch.empty((K, n), dtype=torch.float64, device=dev)
    ctx.barrier()
    t0 = time.perf_counter()
    Vd.copy_(Vh_p, non_blocking=True)
    Z = fac.update_many(Vd, check_diag=False, p2p=p2p, fuse=fuse)
    Zh.copy_(Z, non_blocking=True)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    out["e2e"] = {"value": K / dt, "unit": "updates/s", "h2d_bytes_per_step": K * n * 8, "d2h_bytes_per_step": K * n * 8,
                  "api": "RowCyclicFactor.update_many(V from pinned host memory; scratch vectors copied back)"}
    fac.close()
    del fac, A, V_bank
    torch.cuda.empty_cache()
    return out


def leg_rowcyclic_fullcheck(ctx, args, n=N_BIG, K=4):
    """A full factor through the same multi-rank code path (peer-memory exchange when N > 1), every
    entry compared with the oracle."""
    from cholupdates_b200.distributed import RowCyclicFactor, rc_owned_rows

    torch = ctx.torch
    p2p = None if args.exchange == "auto" else (args.exchange == "p2p")
    seed0 = 70000
    A = rc_make_local(ctx, n, seed0)
    full0 = rc_gather_leading(ctx, A, n, n)
    fac = RowCyclicFactor(A, n, ctx.rank, ctx.world)
    gv = torch.Generator(device=ctx.dev)
    gv.manual_seed(11)
    V = torch.randn((K, n), dtype=torch.float64, device=ctx.dev, generator=gv)
    fac.update_many(V, check_diag=True, p2p=p2p)
    full1 = rc_gather_leading(ctx, A, n, n)
    out = None
    if ctx.rank == 0:
        Vh = V.cpu().numpy()
        ref = oracle_apply(full0.cpu().numpy(), [("update", Vh[u]) for u in range(K)])
        out = parity_report(full1.cpu().numpy(), ref)
        out["what"] = f"all of an n={n} factor after update_many(K={K}) on {ctx.world} rank(s) vs the oracle"
    fac.close()
    del fac, A, full0, full1
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------- leg: the public API as a caller sees it
def leg_public_api(ctx, args):
    torch, rank_1 = ctx.torch, ctx.rank_1
    dev = ctx.dev
    out = {}
    # CUDA tensors through rank_1.update (includes the workspace lookup and the 4-byte status read)
    for n in (1000, N_BIG):
        L = make_factor_cuda(torch, n, 5, dev)
        V = torch.randn((12, n), dtype=torch.float64, device=dev)
        for i in range(3):
            rank_1.update(L, V[i], check_diag=False, overwrite_L=True, overwrite_v=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3, 11):
            rank_1.update(L, V[i], check_diag=False, overwrite_L=True, overwrite_v=True)
        torch.cuda.synchronize()
        out[f"cuda_tensor_n{n}_ms_per_update"] = 1e3 * (time.perf_counter() - t0) / 8
        del L
    if ctx.rank == 0 and not args.no_cpu:
        # configs[0]: the notebook workload, NumPy in / NumPy out, next to the reference on the host
        mod, kind = load_cpu_reference()
        for n in (1000, 5000):
            L0, v0 = config1_inputs(n)
            row = {}
            for order in ("C", "F"):
                Lo = np.asarray(L0, order=order)
                for mode, kw in (("copy", {}), ("inplace", {"overwrite_L": True, "overwrite_v": True})):
                    # (steady state of a caller's loop: after a compute-only stretch the first transfers run on a
                    # PCIe link that is still in its low-power state -- 3 ms instead of 0.6 ms at n = 1000)
                    t_warm = time.perf_counter()
                    while time.perf_counter() - t_warm < 0.15:
                        rank_1.update(Lo.copy(order="K"), v0.copy(), check_diag=False, **kw)
                    ts = []
                    for _ in range(6):
                        Lx, vx = Lo.copy(order="K"), v0.copy()
                        t0 = time.perf_counter()
                        res = rank_1.update(Lx, vx, check_diag=False, **kw)
                        ts.append(time.perf_counter() - t0)
                    row[f"b200_{order}_{mode}_ms"] = 1e3 * min(ts[1:])
                if order == "F":
                    ref = np.array(Lo, order="F", copy=True)
                    mod.update(ref, v0.copy())
                    row["parity"] = parity_report(res, ref)
                ts = []
                for _ in range(4):
                    Lx, vx = Lo.copy(order="K"), v0.copy()
                    t0 = time.perf_counter()
                    mod.update(Lx, vx)
                    ts.append(time.perf_counter() - t0)
                row[f"reference_{order}_inplace_ms"] = 1e3 * min(ts[1:])
            row["reference_kind"] = kind
            out[f"numpy_config1_n{n}"] = row
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------- GPU arm
def safe_leg(ctx, name, fn):
    """Run one extra leg; a failure is recorded, not fatal (the headline line must still be printed)."""
    t0 = time.perf_counter()
    try:
        res = fn()
    except Exception as exc:  # noqa: BLE001
        traceback.print_exc(file=sys.stderr)
        res = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        try:
            ctx.torch.cuda.empty_cache()
        except Exception:  # noqa: BLE001
            pass
    if isinstance(res, dict):
        res["leg_wall_s"] = round(time.perf_counter() - t0, 2)
    return res


def run_gpu_arm(args):
    ctx = Ctx(args)
    sampler = ClockSampler(ctx.local)
    sampler.start()  # early: the first NVML queries are slow, and the timed region is short
    wl = args.workload
    line_extra = {}
    if wl in ("all", "update"):
        top = leg_update(ctx, args, sampler)
        if wl == "all":
            legs = {}
            legs["downdate_n16384"] = safe_leg(ctx, "downdate", lambda: leg_downdate(ctx, args))
            other = safe_leg(ctx, "layouts_and_sizes", lambda: leg_layouts_and_sizes(ctx, args))
            if isinstance(other, dict) and "error" not in other:
                other.pop("leg_wall_s", None)
                legs.update(other)
            else:
                legs["layouts_and_sizes"] = other
            legs["batched_8192x256"] = safe_leg(ctx, "batched", lambda: leg_batched(ctx, args))
            rc = safe_leg(ctx, "rowcyclic", lambda: leg_rowcyclic(ctx, args, RC_N, RC_K))
            if isinstance(rc, dict) and "error" not in rc:
                rc["parity_full_n16384"] = safe_leg(ctx, "rowcyclic_full", lambda: leg_rowcyclic_fullcheck(ctx, args))
            legs["rowcyclic_n65536_k64"] = rc
            legs["public_api"] = safe_leg(ctx, "public_api", lambda: leg_public_api(ctx, args))
            top["roofline"]["workloads"] = legs
    elif wl == "downdate":
        sampler.stop()
        d = leg_downdate(ctx, args)
        top = {"metric": "rank-1 downdates/s, n=16384 fp64 in-place", "value": d["value"], "unit": d["unit"],
               "ms_per_step": d["ms_per_downdate"], "scaling": "weak", "config": {"workload": d["workload"]},
               "roofline": dict(d["roofline"], parity=d.get("parity"), not_pd_case=d.get("not_pd_case")),
               "cpu_baseline": None, "e2e": None, "gpu_launches": None, "clocks": None}
        args.steps, args.warmup = d["steps"], d["warmup"]
    elif wl == "batched":
        sampler.stop()
        d = leg_batched(ctx, args)
        top = {"metric": METRIC_BATCHED, "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_batch"],
               "scaling": "weak", "config": {"workload": d["workload"], "batch_per_gpu": BATCH, "n": N_SMALL,
                                             "parallelism": d["parallelism"]},
               "roofline": dict(d["roofline"], parity=d.get("parity")), "cpu_baseline": None, "e2e": None,
               "gpu_launches": None, "clocks": None}
        args.steps, args.warmup = d["steps"], d["warmup"]
    else:  # rowcyclic
        sampler.stop()
        d = leg_rowcyclic(ctx, args)
        top = {"metric": f"successive rank-1 updates/s of one n={d['n']} fp64 factor, block-row-cyclic over the job's GPUs "
                         f"({d['updates_per_step']} per step)",
               "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "ms_per_update": d["ms_per_update"],
               "scaling": "strong", "config": {"workload": d["workload"], "exchange": d["exchange"],
                                                "parallelism": d["parallelism"], "l2": "inputs_larger_than_L2"},
               "roofline": dict(d["roofline"], parity=d.get("parity")), "cpu_baseline": None, "e2e": d.get("e2e"),
               "gpu_launches": d["gpu_launches"], "clocks": None}
        args.steps, args.warmup = d["steps"], d["warmup"]
    if ctx.world > 1:
        ctx.dist.barrier()
    if ctx.rank == 0:
        line = {"metric": top["metric"], "value": top["value"], "unit": top["unit"], "n_gpus": ctx.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": top["ms_per_step"],
                "higher_is_better": True, "scaling": top["scaling"], "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": top["config"], "roofline": top["roofline"],
                "cpu_baseline": top["cpu_baseline"], "e2e": top["e2e"], "gpu_launches": top["gpu_launches"],
                "clocks": top["clocks"]}
        for k in ("ms_per_update", "step_ms"):
            if k in top:
                line[k] = top[k]
        line.update(line_extra)
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-worker"])
    ap.add_argument("--workload", default="all", choices=["all", "update", "downdate", "batched", "rowcyclic"])
    ap.add_argument("--rc-n", dest="n", type=int, default=0, help="rowcyclic: order of the factor (default 65536)")
    ap.add_argument("--rc-k", dest="k", type=int, default=RC_K, help="rowcyclic: successive updates per step")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="rowcyclic: coefficient exchange")
    ap.add_argument("--rc-fuse", dest="fuse", type=int, default=0, help="rowcyclic: updates per pass over a panel (0 = auto)")
    ap.add_argument("--layout", default="C", choices=["C", "F"])
    ap.add_argument("--path", type=int, default=0, help="force kernel path (1 small, 2 panel, 3 dataflow)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / reference timings")
    ap.add_argument("--quick", action="store_true", help="headline leg: skip the pageable / resident e2e variants")
    ap.add_argument("--seed", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--sync-file", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference-worker":
        reference_worker(args)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
