// rowcyclic.cuh -- one huge factor distributed block-row-cyclically over G GPUs.
//
// Rank g owns the global 256-row blocks R = g, g + G, g + 2G, ... and stores them back to back as a
// local row-major matrix A_loc[m_loc][ld] (all n columns).  Row i only ever needs the coefficients of
// columns k < i plus its own data, so the only exchange per 256-column panel J is the
// panel's coefficients (p, c, sigma, L'_kk and the running t, S): ~8 KB broadcast from the owner of
// the diagonal block, rank J mod G.  Per panel:
//     owner:  diagonal inverses + chain (large.cuh kernels on its local rows) -> pack
//     all:    broadcast (NCCL / torch.distributed, done by the host driver) -> unpack
//     all:    trailing update of the local rows below the panel (this file)
// This is the launch-per-panel formulation; fusing it into the persistent dataflow kernel with a
// pipelined exchange is the next step.
#pragma once

#include "dataflow.cuh"
#include "large.cuh"

namespace chub {

constexpr int kRcBlock = kBig;              // rows per distribution block (= panel width)
constexpr int kRcPayload = 4 * kBig + 8;    // doubles per panel message: p, c, sigma, dn, scal[0..7]

__device__ __forceinline__ int64_t rc_global_row(int64_t l, int G, int rank) {
  return ((l / kRcBlock) * G + rank) * kRcBlock + (l % kRcBlock);
}

// w <- v for the owned rows, check_diag on the owned diagonal entries, t <- 1, S <- 1.
template <typename T>
__global__ void rc_prepare_kernel(const T *A, int m_loc, int n, int64_t ld, int G, int rank, const T *v,
                                  LargeWs ws, int flags, int32_t *status) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < m_loc) {
    const int64_t R = rc_global_row(l, G, rank);
    if (R < n) {
      ws.w[R] = (double)v[R];
      if ((flags & CHUB_FLAG_CHECK_DIAG) && A[(int64_t)l * ld + R] == T(0))
        atomicMax(status, CHUB_STATUS_ZERO_DIAG);
    }
  }
  if (l == 0) { ws.scal[0] = 1.0; ws.scal[1] = 1.0; ws.scal[2] = 0.0; ws.scal[3] = 0.0; }
}

// (n_pad = length of the workspace vectors: a ragged last panel must not reach past them)
__global__ void rc_pack_kernel(LargeWs ws, int J, int n_pad, double *buf, const int32_t *status) {
  const int t = threadIdx.x;  // kBig threads
  const int k = J * kBig + t;
  const bool in = k < n_pad;
  buf[t] = in ? ws.p[k] : 0.0;
  buf[kBig + t] = in ? ws.c[k] : 1.0;
  buf[2 * kBig + t] = in ? ws.sg[k] : 0.0;
  buf[3 * kBig + t] = in ? ws.dn[k] : 1.0;
  if (t < 7) buf[4 * kBig + t] = ws.scal[t];
  if (t == 7) buf[4 * kBig + 7] = (double)*status;  // the owner's status travels with the panel
}

__global__ void rc_unpack_kernel(LargeWs ws, int J, int n_pad, const double *buf, int32_t *status) {
  const int t = threadIdx.x;
  const int k = J * kBig + t;
  if (k < n_pad) {
    ws.p[k] = buf[t];
    ws.c[k] = buf[kBig + t];
    ws.sg[k] = buf[2 * kBig + t];
    ws.dn[k] = buf[3 * kBig + t];
  }
  if (t < 7) ws.scal[t] = buf[4 * kBig + t];
  if (t == 7 && buf[4 * kBig + 7] != 0.0) atomicMax(status, (int)buf[4 * kBig + 7]);
}

// Trailing update of the local rows whose global index is >= J*256 (same arithmetic as
// large_trail_kernel<T, ROW_MAJOR, true>; one CTA = 128 consecutive local = global rows).
// UPDATE = false is the read-only pass of the downdate's forward substitution (w <- w - L p on the
// rows below the panel's diagonal block; the block's own rows are finished by the chain).
template <typename T, bool UPDATE = true>
__global__ void __launch_bounds__(128) rc_trail_kernel(T *A, int m_loc, int n, int64_t ld, int G, int rank,
                                                       LargeWs ws, int J, const int32_t *status) {
  constexpr int NT = 128;
  __shared__ double cs[3][kPanel];
  __shared__ __align__(16) T tile[NT * kPanel];
  if (*status != 0) return;
  const int tid = threadIdx.x;
  const int l0 = blockIdx.x * NT;
  if (l0 >= m_loc) return;
  const int64_t g0 = rc_global_row(l0, G, rank);
  const int base = J * kBig;
  int rows = min(NT, m_loc - l0);
  if (g0 + rows > n) rows = (int)max((int64_t)0, (int64_t)n - g0);
  if (rows <= 0 || g0 + rows <= base + (UPDATE ? 0 : kBig)) return;  // entirely above the panel
  const int r0 = (int)g0;
  T *L = A + ((int64_t)l0 - g0) * ld;  // so that L[R * ld + c] addresses global row R
  const int R = r0 + tid;
  const bool valid = tid < rows && R >= base + (UPDATE ? 0 : kBig);
  double w = valid ? ws.w[R] : 0.0;
  const int cend = min(n, base + kBig);
  for (int c0 = base; c0 < cend; c0 += kPanel) {
    const int bw = min(kPanel, n - c0);
    if (c0 > r0 + rows - 1) break;
    __syncthreads();
    if (tid < kPanel) {
      cs[0][tid] = tid < bw ? ws.p[c0 + tid] : 0.0;
      if (UPDATE) {
        cs[1][tid] = tid < bw ? ws.c[c0 + tid] : 1.0;
        cs[2][tid] = tid < bw ? ws.sg[c0 + tid] : 0.0;
      }
    }
    tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
    __syncthreads();
    if (valid) {
#pragma unroll 8
      for (int k = 0; k < kPanel; ++k) {
        if (k < bw && c0 + k < R) {
          const int o = tile_off(tid, k);
          const double lo = (double)tile[o];
          if (UPDATE) tile[o] = (T)(cs[1][k] * lo + cs[2][k] * w);
          w -= lo * cs[0][k];
        }
      }
      if (UPDATE && R >= c0 && R < c0 + bw) tile[tile_off(tid, R - c0)] = (T)ws.dn[R];
    }
    if (UPDATE) {
      __syncthreads();
      // rows above the panel inside this CTA (R < base) are loaded but never stored back changed
      tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
    }
  }
  if (valid && R >= cend) ws.w[R] = w;
}

// ---- downdate: distributed forward substitution, then a purely local sweep
// Owner of panel J: p of the panel (from the chain) and the signs of its input diagonal entries.
template <typename T>
__global__ void rc_pack_dd_kernel(const T *Lsh, int n, int64_t ld, LargeWs ws, int J, double *buf,
                                  const int32_t *status) {
  const int t = threadIdx.x;  // kBig threads
  const int k = J * kBig + t;
  buf[t] = k < n ? ws.p[k] : 0.0;
  buf[3 * kBig + t] = (k < n && Lsh[(int64_t)k * ld + k] < T(0)) ? -1.0 : 1.0;
  if (t == 7) buf[4 * kBig + 7] = (double)*status;
}
__global__ void rc_unpack_dd_kernel(LargeWs ws, int J, int n_pad, const double *buf, int32_t *status) {
  const int t = threadIdx.x;
  const int k = J * kBig + t;
  if (k < n_pad) {
    ws.p[k] = buf[t];
    ws.dn[k] = buf[3 * kBig + t];
  }
  if (t == 7 && buf[4 * kBig + 7] != 0.0) atomicMax(status, (int)buf[4 * kBig + 7]);
}

// Reverse sweep of the local rows (same arithmetic as large_dd_sweep_kernel<T, ROW_MAJOR>):
// rows are independent once every rank holds all (c_k, s_k, sign_k), so there is no exchange.
// One CTA = 128 consecutive local = global rows.
template <typename T>
__global__ void __launch_bounds__(128) rc_dd_sweep_kernel(T *A, int m_loc, int n, int64_t ld, int G, int rank,
                                                          LargeWs ws, const int32_t *status) {
  constexpr int NT = 128;
  __shared__ double cs[3][kPanel];
  __shared__ __align__(16) T tile[NT * kPanel];
  if (*status != 0) return;
  const int tid = threadIdx.x;
  const int nblk = (m_loc + NT - 1) / NT;
  const int l0 = (nblk - 1 - (int)blockIdx.x) * NT;  // longest rows first
  const int64_t g0 = rc_global_row(l0, G, rank);
  int rows = min(NT, m_loc - l0);
  if (g0 + rows > n) rows = (int)max((int64_t)0, (int64_t)n - g0);
  if (rows <= 0) return;
  const int r0 = (int)g0;
  T *L = A + ((int64_t)l0 - g0) * ld;
  const int R = r0 + tid;
  const bool valid = tid < rows;
  double x = 0.0;
  for (int c0 = ((r0 + rows - 1) / kPanel) * kPanel; c0 >= 0; c0 -= kPanel) {
    const int bw = min(kPanel, n - c0);
    __syncthreads();
    if (tid < kPanel) {
      cs[0][tid] = tid < bw ? ws.c[c0 + tid] : 1.0;
      cs[1][tid] = tid < bw ? ws.sg[c0 + tid] : 0.0;
      cs[2][tid] = tid < bw ? ws.dn[c0 + tid] : 1.0;
    }
    tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
    __syncthreads();
    if (valid) {
#pragma unroll 8
      for (int kk = kPanel - 1; kk >= 0; --kk) {
        if (kk < bw && c0 + kk <= R) {
          const int o = tile_off(tid, kk);
          const double lo = (double)tile[o], c = cs[0][kk], s = cs[1][kk];
          tile[o] = (T)(cs[2][kk] * (c * lo - s * x));
          x = c * x + s * lo;
        }
      }
    }
    __syncthreads();
    tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
  }
}

// =====================================================================================
// Wavefront over K successive updates (BASELINE.json configs[4]: "64 successive updates").
//
// Task (u, J) = panel J of update u.  It needs task (u-1, J) (the panel after the previous update)
// and task (u, J-1) (the residual w^{(u)} entering the panel), so all tasks on an anti-diagonal
// d = u + J are independent: they touch different column panels and different w vectors.  One step =
// one anti-diagonal:
//     owner kernel  -- every rank, one CTA per task it owns (J mod G == rank): inverse diagonal
//                      blocks + chain -> the task's payload (p, c, sigma, L'_kk, z, {t, S})
//     all-gather    -- ONE exchange per step (host driver, NCCL): G * max_tasks payloads
//     trail kernel  -- every rank, grid (row blocks, tasks of the step): applies each panel to the
//                      local rows below it
// K + nJ - 1 steps replace K * nJ per-panel broadcasts; the per-step latency is amortised over up
// to min(K, nJ) panels of streaming work.  Same arithmetic as the per-panel kernels above.
constexpr size_t kRcwOwnerSmem = (size_t)(kBig / kPanel) * kInvStageDoubles * sizeof(double);  // 69 632 B
constexpr int kRcwTrailNT = 128;           // rows per CTA of the trail kernel (32 / 64 / 128)
constexpr int kRcwPayload = 5 * kBig + 8;  // p, c, sigma, dn, z, scal[8] ({t, S} leaving the panel)

struct RcWave {
  double *W;     // [K][n_pad]  residual vector of every update (entries of the owned rows are live)
  double *scal;  // [K][8]      {t, S} of every update entering its next panel
  double *X;     // [max_tasks][8][kXBlock]  inverse diagonal blocks of this rank's tasks of the step
  double *D;     // [max_tasks][256][256]    fused mode: working copy of a task's diagonal block
  int64_t n_pad;
  int K;
  int F;         // updates per task (1, or kRcwMaxFuse: consecutive updates share one pass over the panel)
};
constexpr int kRcwMaxFuse = 2;
// Fused mode: "update" in the task grid means a GROUP of F consecutive updates; NU groups in all.
__host__ __device__ inline int64_t rcw_groups(int64_t K, int F) { return (K + F - 1) / F; }

inline int rcw_panels(int64_t n) { return (int)((n + kBig - 1) / kBig); }
inline int rcw_max_tasks(int64_t n, int64_t K, int G) {
  const int64_t len = K < rcw_panels(n) ? K : rcw_panels(n);
  return (int)((len + G - 1) / G) > 0 ? (int)((len + G - 1) / G) : 1;
}
inline size_t rcw_ws_doubles(int64_t n, int64_t K, int G) {
  const int64_t n_pad = (n + 31) & ~31LL;
  return (size_t)(K * n_pad + K * 8 + (int64_t)rcw_max_tasks(n, K, G) * ((kBig / kPanel) * kXBlock + kBig * kBig));
}
inline RcWave rcw_carve(void *base, int64_t n, int64_t K, int G) {
  RcWave wv;
  wv.n_pad = (n + 31) & ~31LL;
  wv.K = (int)K;
  double *d = reinterpret_cast<double *>(base);
  wv.W = d; d += K * wv.n_pad;
  wv.scal = d; d += K * 8;
  wv.X = d; d += (int64_t)rcw_max_tasks(n, K, G) * (kBig / kPanel) * kXBlock;
  wv.D = d;
  wv.F = 1;
  return wv;
}

// ---- exchange over NVLink peer memory instead of NCCL (one process per GPU, cudaIpc-mapped mailboxes).
// Every rank owns a MAILBOX: for each parity q of the global step count and each sender rank r, max_tasks payload
// slots plus one flag.  The owner kernel of step d stores its payloads straight into every peer's
// mailbox (plain stores over NVLink), and its last CTA then raises flag[q][rank] = epoch + d + 1 on
// every peer (fence.sys, then the flag: payload before flag).  The trail kernel's CTAs spin (acquire) on
// the flag of the task's owner in their OWN mailbox before they read the payload.  Every rank raises
// its flag on every peer at EVERY step (also when it owns no task), and one CTA of every trail kernel
// waits for all G flags: no rank can be more than one step ahead of another, which is what makes two
// parities enough.  No NCCL call, no host synchronisation between steps: 2 launches per step.
constexpr int kRcwSets = 4;
constexpr int kRcwSlot = 2 * kRcwPayload;  // a task's slot holds the payloads of up to kRcwMaxFuse updates
struct RcwP2P {
  double *const *peer_mail;  // device array [G]: mailbox base of every rank as mapped here (nullptr = NCCL mode)
  double *mail;              // this rank's own mailbox
  long long epoch;           // global step of this call's step 0 (monotonic across calls): step d has
                             // parity (epoch + d) & 1 and flag value epoch + d + 1
  int max_tasks;             // slot stride of the MAILBOX: rcw_mail_tasks(n, G), the same for every call on it
                             // (a per-call stride would let calls with different K alias each other's slots)
  long long watchdog;        // cycles a CTA may wait for a peer's flag before it gives up
};
// Tasks per (set, sender) the mailbox has room for: the most any call can have, ceil(panels / G).
inline int rcw_mail_tasks(int64_t n, int G) { return rcw_max_tasks(n, rcw_panels(n), G); }
// Mailbox layout (doubles): [ long long flag[4][G] | int counter, padding (8) | payload slots [4][G][max_tasks] ].
// Four sets (global step & 3): with the look-ahead split the bulk of step d's trail may still be reading
// its payloads while the payloads of step d+2 arrive (never those of step d+3, see distributed.py).
// The flags sit at a fixed place, so calls with different K (different max_tasks) can share a mailbox
// that was sized for the largest one.
inline size_t rcw_mailbox_doubles(int64_t n, int64_t K, int G) {
  (void)K;  // sized for every K: the slot stride is fixed per mailbox
  return (size_t)kRcwSets * G + 8 + (size_t)kRcwSets * G * rcw_mail_tasks(n, G) * kRcwSlot;
}
__host__ __device__ inline size_t rcw_mail_slot(int q, int r, int i, int G, int max_tasks) {
  return (size_t)kRcwSets * G + 8 + ((size_t)(q * G + r) * max_tasks + i) * kRcwSlot;
}
__host__ __device__ inline size_t rcw_mail_flags(int G, int max_tasks) {
  (void)G; (void)max_tasks;
  return 0;  // long long flag[kRcwSets][G], then an int counter
}
__device__ __forceinline__ void st_release_sys_s64(long long *p, long long x) {
  asm volatile("st.release.sys.global.s64 [%0], %1;\n" ::"l"(p), "l"(x) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys_s64(const long long *p) {
  long long x;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];\n" : "=l"(x) : "l"(p) : "memory");
  return x;
}
// Spin until *flag >= want (thread 0 of a CTA); watchdog -> CHUB_STATUS_INTERNAL instead of a hang.
// After CHUB_STATUS_INTERNAL nobody stores to A any more (the callers bail out), the factor of this call
// is invalid on every rank and the mailbox has to be re-created (distributed.py raises).
__device__ __forceinline__ void rcw_wait_flag(const long long *flag, long long want, int32_t *status,
                                              long long watchdog) {
  if (ld_acquire_sys_s64(flag) >= want) return;
  const long long t0 = clock64();
  while (ld_acquire_sys_s64(flag) < want) {
    if (clock64() - t0 > watchdog || *(volatile int32_t *)status == CHUB_STATUS_INTERNAL) {
      atomicMax(status, CHUB_STATUS_INTERNAL);
      return;
    }
  }
}

// W[u] <- V[u] on the owned rows, check_diag on the owned diagonal entries, {t, S} <- {1, 1}.
template <typename T>
__global__ void rcw_prepare_kernel(const T *A, int m_loc, int n, int64_t ld, int G, int rank, const T *V,
                                   int64_t ldv, RcWave wv, int flags, int32_t *status) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int u = blockIdx.y;
  if (l < m_loc) {
    const int64_t R = rc_global_row(l, G, rank);
    if (R < n) {
      wv.W[(int64_t)u * wv.n_pad + R] = (double)V[(int64_t)u * ldv + R];
      if (u == 0 && (flags & CHUB_FLAG_CHECK_DIAG) && A[(int64_t)l * ld + R] == T(0))
        atomicMax(status, CHUB_STATUS_ZERO_DIAG);
    }
  }
  if (l == 0) {
    wv.scal[u * 8 + 0] = 1.0;
    wv.scal[u * 8 + 1] = 1.0;
  }
}

// One CTA (256 threads) per task of this rank on anti-diagonal d: J = Jfirst + blockIdx.x * G,
// u = d - J.  Writes the task's payload into pay_mine[blockIdx.x].
// P2P mode (pp.peer_mail != nullptr): `count` = number of tasks (0: this CTA only raises the flags).
// UPDATE = false: forward substitution only (the downdate's p = L^{-1} v): the payload carries p and the
// signs of the panel's input diagonal entries (slot of L'_kk).
template <typename T, bool UPDATE = true>
__global__ void __launch_bounds__(kBig) rcw_owner_kernel(const T *A, int n, int64_t ld, int G, int rank, RcWave wv,
                                                         int d, int Jfirst, double *pay_mine, RcwP2P pp, int count,
                                                         int32_t *status) {
  if (*status != 0) return;
  const int i = blockIdx.x;
  if (pp.peer_mail && count == 0) {  // nothing to compute at this step: heartbeat only
    if (threadIdx.x < G)
      st_release_sys_s64(reinterpret_cast<long long *>(pp.peer_mail[threadIdx.x] + rcw_mail_flags(G, pp.max_tasks)) +
                             (int)((pp.epoch + d) & (kRcwSets - 1)) * G + rank, pp.epoch + d + 1);
    return;
  }
  const int J = Jfirst + i * G, U = d - J;  // U = update (F == 1) or group of F consecutive updates
  const int F = UPDATE ? wv.F : 1;
  const int Fe = min(F, wv.K - U * F);       // updates in this group (the last group may be short)
  // the diagonal block of panel J is local block J / G; shift the base so that GLOBAL row indices work
  const T *Lsh = A + ((int64_t)(J / G) - (int64_t)J) * kRcBlock * ld;
  double *X = wv.X + (int64_t)i * (kBig / kPanel) * kXBlock;
  const int warp = threadIdx.x >> 5;
  extern __shared__ __align__(16) double rcw_stage[];  // [8][kInvStageDoubles]
  double *stg = rcw_stage + (size_t)warp * kInvStageDoubles;
  double *pay = pay_mine + (int64_t)i * kRcwSlot;
  if (F == 1) {
    // the 8 inverse diagonal blocks at once, one warp each (staging in dynamic shared memory)
    invdiag_block_smem<T, CHUB_ROW_MAJOR>(Lsh, n, ld, X + (int64_t)warp * kXBlock, J * (kBig / kPanel) + warp,
                                          reinterpret_cast<double (*)[kPanel + 1]>(stg), stg + kPanel * (kPanel + 1));
    __syncthreads();
    chain_body<T, CHUB_ROW_MAJOR, UPDATE>(Lsh, n, ld, J, X, wv.W + (int64_t)U * wv.n_pad, pay, pay + kBig,
                                          pay + 2 * kBig, pay + 3 * kBig, (T *)nullptr, pay + 4 * kBig,
                                          wv.scal + U * 8, pay + 5 * kBig);
    if (!UPDATE) {
      const int R = J * kBig + (int)threadIdx.x;
      pay[3 * kBig + threadIdx.x] = (R < n && Lsh[(int64_t)R * ld + R] < T(0)) ? -1.0 : 1.0;
    }
  } else {
    // Fused: the coefficients of update f+1 need the diagonal block AFTER update f.  The block is copied
    // into a scratch (fp64, ld = 256), and between two chains update f is applied to the scratch by the
    // thread that owns the row.  A itself is left alone: the trail kernel applies all Fe updates to the
    // block's rows together with every other row (the same arithmetic, from the same payloads).
    const int base = J * kBig;
    double *Dm = wv.D + (int64_t)i * kBig * kBig;
    const double *Dsh = Dm - ((int64_t)base * kBig + base);  // Dsh[R * 256 + C] for global R, C of the block
    for (int r = warp; r < kBig; r += kBig / 32) {
      const int R = base + r;
      if (R < n)
        for (int c = threadIdx.x & 31; c <= r; c += 32) Dm[r * kBig + c] = (double)Lsh[(int64_t)R * ld + base + c];
    }
    __syncthreads();
    __shared__ double cf[4][kBig];  // p, c, sigma, dn of the update being applied to the scratch
    for (int f = 0; f < Fe; ++f) {
      const int u = U * F + f;
      double *pf = pay + (int64_t)f * kRcwPayload;
      invdiag_block_smem<double, CHUB_ROW_MAJOR>(Dsh, n, kBig, X + (int64_t)warp * kXBlock, J * (kBig / kPanel) + warp,
                                                 reinterpret_cast<double (*)[kPanel + 1]>(stg), stg + kPanel * (kPanel + 1));
      __syncthreads();
      chain_body<double, CHUB_ROW_MAJOR, true>(Dsh, n, kBig, J, X, wv.W + (int64_t)u * wv.n_pad, pf, pf + kBig,
                                               pf + 2 * kBig, pf + 3 * kBig, (double *)nullptr, pf + 4 * kBig,
                                               wv.scal + u * 8, pf + 5 * kBig);
      __syncthreads();
      if (f + 1 < Fe) {
        const int a = threadIdx.x, R = base + a;
#pragma unroll
        for (int q = 0; q < 4; ++q) cf[q][a] = pf[q * kBig + a];
        __syncthreads();
        if (R < n) {
          double w = wv.W[(int64_t)u * wv.n_pad + R];
          double *row = Dm + a * kBig;
#pragma unroll 8
          for (int c = 0; c < a; ++c) {
            const double lo = row[c];
            row[c] = fma(cf[1][c], lo, cf[2][c] * w);
            w = fma(-lo, cf[0][c], w);
          }
          row[a] = cf[3][a];
        }
        __syncthreads();
      }
    }
  }
  if (!pp.peer_mail) return;
  // ---- push the payload into every rank's mailbox (own included), then the flags
  __syncthreads();
  const int q = (int)((pp.epoch + d) & (kRcwSets - 1));
  for (int g = 0; g < G; ++g) {
    double *dst = pp.peer_mail[g] + rcw_mail_slot(q, rank, i, G, pp.max_tasks);
    for (int e = threadIdx.x; e < Fe * kRcwPayload; e += kBig) dst[e] = pay[e];
  }
  __threadfence_system();  // my stores are performed system-wide before I am counted as done
  __syncthreads();
  __shared__ int last;
  int *counter = reinterpret_cast<int *>(reinterpret_cast<long long *>(pp.mail + rcw_mail_flags(G, pp.max_tasks)) + kRcwSets * G);
  if (threadIdx.x == 0) {
    const int done = atomicAdd(counter, 1);
    last = (done == count - 1);
    if (last) *counter = 0;  // (the next owner kernel of this rank starts after this one has ended)
  }
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < G)
      st_release_sys_s64(reinterpret_cast<long long *>(pp.peer_mail[threadIdx.x] + rcw_mail_flags(G, pp.max_tasks)) +
                             q * G + rank, pp.epoch + d + 1);
  }
}

// Trailing update for every task of anti-diagonal d: blockIdx.y = t <-> J = Jlo + t, u = d - J; its
// payload sits in slot (J mod G) * max_tasks + t / G of the gathered buffer (the owner's tasks are
// Jfirst, Jfirst + G, ... with Jfirst - Jlo < G).  blockIdx.x = NT consecutive local rows.
// UPDATE = false: the read-only pass of the downdate's forward substitution (w <- w - L p on the rows
// below the panel's diagonal block); the unpacking CTA stores p and the diagonal signs of the panel into
// the n-vectors p_out / dn_out (what the rho^2 test and the sweep read) instead of {t, S} and V_out.
template <typename T, int NT, bool UPDATE = true, int F = 1>
__global__ void __launch_bounds__(NT) rcw_trail_kernel(T *A, int m_loc, int n, int64_t ld, int G, int rank, RcWave wv,
                                                       int d, int Jlo, int max_tasks, const double *pay_all,
                                                       T *V_out, int64_t ldv, RcwP2P pp, int rows_mode,
                                                       double *p_out, double *dn_out, int32_t *status) {
  __shared__ __align__(16) double cp[F][kPanel];    // p_k of each of the task's updates
  __shared__ __align__(16) double2 ccs[F][kPanel];  // (c_k, sigma_k): one 16-byte load per column
  __shared__ __align__(16) T tile[NT * kPanel];
  if (*status != 0) return;
  const int tid = threadIdx.x;
  const int J = Jlo + (int)blockIdx.y, U = d - J;  // U = update (F == 1) or group of F consecutive updates
  const int Fe = min(F, wv.K - U * F);
  const double *pay = pay_all + ((int64_t)(J % G) * max_tasks + blockIdx.y / G) * kRcwSlot;
  const int base = J * kBig;
  // Which local rows does this CTA take?  rows_mode 0: all rows below the panel's top (blockIdx.x = NT
  // consecutive local rows).  Look-ahead split: 1 = HEAD, only the two 256-row blocks J and J+1 (what
  // the owner kernel of the next step needs; blockIdx.x < 2 * 256 / NT), 2 = BULK, blocks >= J+2.
  int l0 = blockIdx.x * NT;
  bool have_rows = l0 < m_loc;
  if (rows_mode == 1) {
    constexpr int per = kRcBlock / NT;  // CTAs per 256-row block
    const int B = J + (int)blockIdx.x / per;
    have_rows = blockIdx.x < 2 * per && B % G == rank && (int64_t)B * kRcBlock < n;
    l0 = (B / G) * kRcBlock + ((int)blockIdx.x % per) * NT;
    have_rows = have_rows && l0 < m_loc;
  }
  int64_t g0 = 0;
  int rows = 0;
  if (have_rows) {
    g0 = rc_global_row(l0, G, rank);
    rows = min(NT, m_loc - l0);
    if (g0 + rows > n) rows = (int)max((int64_t)0, (int64_t)n - g0);
    // entirely above the panel (or, for the bulk, not below the head)?
    have_rows = rows > 0 && g0 + rows > (rows_mode == 2 ? base + 2 * kRcBlock : (UPDATE ? base : base + kRcBlock));
  }
  // CTA x = 0 of every task also unpacks: {t, S} for the next panel of each update (modes 0, 1) and the
  // scratch vectors for the caller (modes 0, 2)
  const bool unpack = blockIdx.x == 0;
  if (!have_rows && !unpack) return;
  if (pp.peer_mail) {
    // P2P mode: the payload arrives in my mailbox.  CTA (0, 0) of the head (or of the unsplit kernel)
    // waits for the flags of ALL ranks -- the one-step bound on how far ranks drift apart.
    const int q = (int)((pp.epoch + d) & (kRcwSets - 1));
    const long long *flags = reinterpret_cast<const long long *>(pp.mail + rcw_mail_flags(G, max_tasks)) + q * G;
    __shared__ int bail;
    if (tid == 0) bail = 0;
    __syncthreads();
    if (blockIdx.x == 0 && blockIdx.y == 0 && rows_mode != 2) {
      if (tid < G) rcw_wait_flag(flags + tid, pp.epoch + d + 1, status, pp.watchdog);
    } else if (tid == 0) {
      rcw_wait_flag(flags + J % G, pp.epoch + d + 1, status, pp.watchdog);
    }
    __syncthreads();
    // a peer never delivered: do not touch A with a stale payload (uniform decision for the CTA)
    if (tid == 0 && *(volatile int32_t *)status == CHUB_STATUS_INTERNAL) bail = 1;
    __syncthreads();
    if (bail) return;
    pay = pp.mail + rcw_mail_slot(q, J % G, blockIdx.y / G, G, pp.max_tasks);
  }
  if (unpack && UPDATE) {
    for (int f = 0; f < Fe; ++f) {
      const double *pf = pay + (int64_t)f * kRcwPayload;
      const int u = U * F + f;
      if (rows_mode != 2 && tid < 2) wv.scal[u * 8 + tid] = __ldcg(pf + 5 * kBig + tid);
      if (rows_mode != 1 && V_out)
        for (int k = tid; k < kBig; k += NT)
          if (base + k < n) V_out[(int64_t)u * ldv + base + k] = (T)__ldcg(pf + 4 * kBig + k);
    }
  }
  if (unpack && !UPDATE && rows_mode != 1)
    for (int k = tid; k < kBig; k += NT)
      if (base + k < n) {
        p_out[base + k] = __ldcg(pay + k);
        dn_out[base + k] = __ldcg(pay + 3 * kBig + k);
      }
  if (!have_rows) return;
  const int r0 = (int)g0;
  T *L = A + ((int64_t)l0 - g0) * ld;  // so that L[R * ld + c] addresses global row R
  const int R = r0 + tid;
  const bool valid = tid < rows && R >= (UPDATE ? base : base + kRcBlock);
  double w[F];
#pragma unroll
  for (int f = 0; f < F; ++f) w[f] = (valid && f < Fe) ? wv.W[(int64_t)(U * F + f) * wv.n_pad + R] : 0.0;
  const int cend = min(n, base + kBig);
  for (int c0 = base; c0 < cend; c0 += kPanel) {
    const int bw = min(kPanel, n - c0);
    if (c0 > r0 + rows - 1) break;
    __syncthreads();
    for (int e = tid; e < F * kPanel; e += NT) {
      const int f = e / kPanel, k = e % kPanel;
      const double *pf = pay + (int64_t)f * kRcwPayload + (c0 - base + k);
      const bool ok = k < bw && f < Fe;
      cp[f][k] = ok ? __ldcg(pf) : 0.0;  // (L2 loads: in P2P mode other GPUs wrote the payload)
      if (UPDATE) ccs[f][k] = make_double2(ok ? __ldcg(pf + kBig) : 1.0, ok ? __ldcg(pf + 2 * kBig) : 0.0);
    }
    tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
    if (tid < rows && c0 + kPanel < cend) {  // next block column of my row -> L2 while this one is processed
      const T *q = L + (int64_t)R * ld + c0 + kPanel;
      prefetch_l2(q);
      if (kPanel * sizeof(T) > 128) prefetch_l2(q + 16);
    }
    __syncthreads();
    if (bw == kPanel && c0 + kPanel <= r0) {
      // the whole tile lies strictly below the diagonal (the common case): no per-entry predicates
      if (tid < rows) {
        T *trow = tile + tid * kPanel;
        const int sw = tid & 31;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          if (f < Fe) {
#pragma unroll
            for (int k = 0; k < kPanel; k += 2) {  // 3 coefficient loads of 16 bytes per 2 columns
              const double2 p2 = *reinterpret_cast<const double2 *>(&cp[f][k]);
              const double l0v = (double)trow[k ^ sw];
              const double l1v = (double)trow[(k + 1) ^ sw];
              if (UPDATE) {
                const double2 q0 = ccs[f][k], q1 = ccs[f][k + 1];
                const double n0 = fma(q0.x, l0v, q0.y * w[f]);
                w[f] = fma(-l0v, p2.x, w[f]);
                const double n1 = fma(q1.x, l1v, q1.y * w[f]);
                w[f] = fma(-l1v, p2.y, w[f]);
                trow[k ^ sw] = (T)n0;
                trow[(k + 1) ^ sw] = (T)n1;
              } else {
                w[f] = fma(-l0v, p2.x, w[f]);
                w[f] = fma(-l1v, p2.y, w[f]);
              }
            }
          }
        }
      }
    } else if (valid) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        if (f < Fe) {
#pragma unroll 8
          for (int k = 0; k < kPanel; ++k) {
            if (k < bw && c0 + k < R) {
              const int o = tile_off(tid, k);
              const double lo = (double)tile[o];
              if (UPDATE) tile[o] = (T)fma(ccs[f][k].x, lo, ccs[f][k].y * w[f]);
              w[f] = fma(-lo, cp[f][k], w[f]);
            }
          }
          if (UPDATE && R >= c0 && R < c0 + bw)
            tile[tile_off(tid, R - c0)] = (T)__ldcg(pay + (int64_t)f * kRcwPayload + 3 * kBig + R - base);
        }
      }
    }
    if (UPDATE) {
      __syncthreads();
      // rows above the panel inside this CTA (R < base) are loaded but stored back unchanged
      tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
    }
  }
  if (valid && R >= cend) {
#pragma unroll
    for (int f = 0; f < F; ++f)
      if (f < Fe) wv.W[(int64_t)(U * F + f) * wv.n_pad + R] = w[f];
  }
}

// ---- bulk of the trail on the TMA ring (look-ahead split, rows_mode 2: the 256-row blocks >= J+2 of every task
// of the step).  These tiles are all full and strictly below the diagonal, and they are where nearly all of a
// step's bytes are.  Same machinery as rect_apply_rm_kernel (dataflow.cuh): persistent warps, a warp takes
// 8 local rows of one task and walks the task's 256 columns in 32-column TMA boxes through a ring that keeps
// running from one (task, row group) item to the next, the block column's (p, c, sigma) riding along as bulk
// copies out of the task's payload; lane mapping and arithmetic of the persistent kernel's workers.  The
// staged kernel above reached 3.5 TB/s with 30 % of its stalls on shared-memory scoreboards and read tiles
// that straddle the diagonal whole; this one reads exactly what it writes.
// Items are dealt round-robin over all warps of the grid: flat index = (groups of earlier tasks) + group.
template <typename T>
__global__ void __launch_bounds__(kRectWarps * 32, 1)
rcw_bulk_tma_kernel(int m_loc, int n, int G, int rank, RcWave wv, int d, int Jlo, int ntasks, int max_tasks,
                    const double *pay_all, T *V_out, int64_t ldv, RcwP2P pp, int32_t *status,
                    const __grid_constant__ CUtensorMap tmapA) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, CPT = RmGeom<T>::CPT;
  constexpr int WTILE = RmGeom<T>::WTILE, STAGE = RectGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char rb_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(rb_smem) + warp * kRectStages;
  unsigned char *mystage = rb_smem + 1024 + (size_t)warp * kRectStages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRectWarps * kRectStages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(rb_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int g = lane >> 3, r8 = lane & 7;
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  const unsigned my_off = hh * 1024 + r8 * 128;
  const int TW = (int)gridDim.x * kRectWarps, gw = (int)blockIdx.x * kRectWarps + warp;
  unsigned it = 0;
  long long base_items = 0;
  bool bail = false;
  for (int t = 0; t < ntasks && !bail; ++t) {
    const int J = Jlo + t, U = d - J;
    // local rows of the owned 256-row blocks with global index >= J + 2
    int qb = J + 2 - rank;
    qb = qb > 0 ? (qb + G - 1) / G : 0;
    const int l_begin = qb * kRcBlock;
    const int groups = l_begin < m_loc ? (m_loc - l_begin + 7) / 8 : 0;
    int first = (int)(((long long)gw - base_items) % TW);
    if (first < 0) first += TW;
    base_items += groups;
    const int c_base = J * kBig;
    const int ncol = min(kBig, n - c_base);
    const int steps_t = (ncol + kPanel - 1) / kPanel;  // (a ragged last panel has no rows below it anyway)
    if (groups == 0 && !(gw == 0)) continue;
    const double *pay = pay_all + ((int64_t)(J % G) * max_tasks + t / G) * kRcwSlot;
    bool have_pay = false;
    auto need_payload = [&]() {  // p2p: the task's payload has arrived in my mailbox (flag of its owner)
      if (have_pay) return;
      if (pp.peer_mail) {
        const int q = (int)((pp.epoch + d) & (kRcwSets - 1));
        const long long *flags = reinterpret_cast<const long long *>(pp.mail + rcw_mail_flags(G, pp.max_tasks)) + q * G;
        if (lane == 0) rcw_wait_flag(flags + J % G, pp.epoch + d + 1, status, pp.watchdog);
        __syncwarp();
        pay = pp.mail + rcw_mail_slot(q, J % G, t / G, G, pp.max_tasks);
        fence_proxy_async_all();  // the acquire above -> the bulk copies below
      }
      have_pay = true;
    };
    // warp 0 of the grid hands the task's scratch vector (rotg's z_k) to the caller, as the staged kernel's
    // unpacking CTA does in this mode
    if (gw == 0 && V_out) {
      need_payload();
      if (*(volatile int32_t *)status != CHUB_STATUS_INTERNAL)
        for (int k = lane; k < kBig; k += 32)
          if (c_base + k < n) V_out[(int64_t)U * ldv + c_base + k] = (T)__ldcg(pay + 4 * kBig + k);
    }
    for (int grp = first; grp < groups; grp += TW) {
      need_payload();
      if (*(volatile int32_t *)status == CHUB_STATUS_INTERNAL) { bail = true; break; }  // a peer never delivered: touch nothing
      const int l0 = l_begin + grp * 8;
      const int l = l0 + r8;
      const int64_t R = l < m_loc ? rc_global_row(l, G, rank) : (int64_t)n;
      double w = R < n ? wv.W[(int64_t)U * wv.n_pad + R] : 0.0;

      auto issue_load = [&](int s) {  // elected lane
        if (s >= steps_t) return;
        const unsigned st = (it + (unsigned)s) % kRectStages;
        void *bar = &full[st];
        unsigned char *stg = mystage + (size_t)st * STAGE;
        const int c0 = c_base + s * kPanel;
        mbar_arrive_expect_tx(bar, (unsigned)WTILE + 768u);
        tma_load_2d(stg, &tmapA, c0, l0, bar);
        if (HALVES == 2) tma_load_2d(stg + 1024, &tmapA, c0 + BOXC, l0, bar);
        bulk_g2s(stg + WTILE, pay + s * kPanel, 256u, bar);
        bulk_g2s(stg + WTILE + 256, pay + kBig + s * kPanel, 256u, bar);
        bulk_g2s(stg + WTILE + 512, pay + 2 * kBig + s * kPanel, 256u, bar);
      };

      if (elect_one()) {
        bulk_wait_read<0>();
        for (int s = 0; s < kRectStages - 1; ++s) issue_load(s);
      }
      __syncwarp();
      for (int s = 0; s < steps_t; ++s) {
        const unsigned st = (it + (unsigned)s) % kRectStages;
        const unsigned ph = ((it + (unsigned)s) / kRectStages) & 1u;
        mbar_wait(&full[st], ph, status);
        unsigned char *stg = mystage + (size_t)st * STAGE;
        unsigned char *rowp = stg + my_off;
        const double *cf = reinterpret_cast<const double *>(stg + WTILE) + g * 8;
        int4 raw[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) raw[i] = *reinterpret_cast<const int4 *>(rowp + (((cb + i) ^ r8) << 4));
        T *x = reinterpret_cast<T *>(raw);
        double lo[8], pre[8];
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(cf + j);
          lo[j] = (double)x[j];
          pre[j] = run;
          run = fma(lo[j], a.x, run);
          lo[j + 1] = (double)x[j + 1];
          pre[j + 1] = run;
          run = fma(lo[j + 1], a.y, run);
        }
        double incl = run;
        double y = __shfl_up_sync(0xffffffffu, incl, 8);
        if (g >= 1) incl += y;
        y = __shfl_up_sync(0xffffffffu, incl, 16);
        if (g >= 2) incl += y;
        const double total = __shfl_sync(0xffffffffu, incl, 24 + r8);
        const double wseg = w - (incl - run);
        w -= total;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 c = *reinterpret_cast<const double2 *>(cf + kPanel + j);
          const double2 sg = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + j);
          x[j] = (T)fma(c.x, lo[j], sg.x * (wseg - pre[j]));
          x[j + 1] = (T)fma(c.y, lo[j + 1], sg.y * (wseg - pre[j + 1]));
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          const int c0 = c_base + s * kPanel;
          tma_store_2d(&tmapA, c0, l0, stg);
          if (HALVES == 2) tma_store_2d(&tmapA, c0 + BOXC, l0, stg + 1024);
          bulk_commit();
          bulk_wait_read<1>();
          issue_load(s + kRectStages - 1);
        }
        __syncwarp();
      }
      it += (unsigned)steps_t;
      if (g == 0 && R < n) wv.W[(int64_t)U * wv.n_pad + R] = w;
    }
  }
  if (elect_one()) bulk_wait_all<0>();
}

// ---- the same with the ring running ACROSS items.  An item is only 8 tiles long, and the kernel above starts
// every one of them with an empty pipeline (a drain, three cold loads, the row's w from global memory): 0.63 vs
// 0.54 ms per update for the staged kernel at n = 16384, K = 32.  Here the warp's items form one stream of
// (item, step) units: the loads run kRectStages - 1 units ahead whatever item they belong to (at most one item
// ahead of the one being consumed), and the next item's w and payload are fetched when the current item starts.
template <typename T>
__global__ void __launch_bounds__(kRectWarps * 32, 1)
rcw_bulk_tma2_kernel(int m_loc, int n, int G, int rank, RcWave wv, int d, int Jlo, int ntasks, int max_tasks,
                     const double *pay_all, T *V_out, int64_t ldv, RcwP2P pp, int32_t *status,
                     const __grid_constant__ CUtensorMap tmapA) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, CPT = RmGeom<T>::CPT;
  constexpr int WTILE = RmGeom<T>::WTILE, STAGE = RectGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char rb_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(rb_smem) + warp * kRectStages;
  unsigned char *mystage = rb_smem + 1024 + (size_t)warp * kRectStages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRectWarps * kRectStages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(rb_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int g = lane >> 3, r8 = lane & 7;
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  const unsigned my_off = hh * 1024 + r8 * 128;
  const int TW = (int)gridDim.x * kRectWarps, gw = (int)blockIdx.x * kRectWarps + warp;

  // the payload of task t (p2p: it arrives in my mailbox; wait for its owner's flag)
  int pay_t = -1;
  const double *pay_p = nullptr;
  bool dead = false;  // a peer never delivered: touch nothing any more
  auto payload = [&](int t) -> const double * {
    if (t == pay_t) return pay_p;
    const int J = Jlo + t;
    pay_p = pay_all + ((int64_t)(J % G) * max_tasks + t / G) * kRcwSlot;
    if (pp.peer_mail) {
      const int q = (int)((pp.epoch + d) & (kRcwSets - 1));
      const long long *flags = reinterpret_cast<const long long *>(pp.mail + rcw_mail_flags(G, pp.max_tasks)) + q * G;
      if (lane == 0) rcw_wait_flag(flags + J % G, pp.epoch + d + 1, status, pp.watchdog);
      __syncwarp();
      pay_p = pp.mail + rcw_mail_slot(q, J % G, t / G, G, pp.max_tasks);
      fence_proxy_async_all();  // the acquire above -> the bulk copies that read the payload
      if (*(volatile int32_t *)status == CHUB_STATUS_INTERNAL) dead = true;
    }
    pay_t = t;
    return pay_p;
  };
  // warp 0 of the grid hands every task's scratch vector (rotg's z_k) to the caller
  if (gw == 0 && V_out) {
    for (int t = 0; t < ntasks; ++t) {
      const double *pay = payload(t);
      if (dead) break;
      const int c_base = (Jlo + t) * kBig, U = d - (Jlo + t);
      for (int k = lane; k < kBig; k += 32)
        if (c_base + k < n) V_out[(int64_t)U * ldv + c_base + k] = (T)__ldcg(pay + 4 * kBig + k);
    }
  }

  struct Item { int l0, c_base, steps, U; const double *pay; double w; int64_t R; bool ok; };
  long long flat = gw, base = 0;  // my next item's flat index; items of the tasks before task tcur
  int tcur = 0;
  auto fetch = [&](Item &im) {
    im.ok = false;
    while (tcur < ntasks && !dead) {
      const int J = Jlo + tcur;
      int qb = J + 2 - rank;  // local rows of the owned 256-row blocks with global index >= J + 2
      qb = qb > 0 ? (qb + G - 1) / G : 0;
      const int l_begin = qb * kRcBlock;
      const int groups = l_begin < m_loc ? (m_loc - l_begin + 7) / 8 : 0;
      if (flat < base + groups) {
        im.l0 = l_begin + (int)(flat - base) * 8;
        im.c_base = J * kBig;
        im.steps = (min(kBig, n - im.c_base) + kPanel - 1) / kPanel;
        im.U = d - J;
        im.pay = payload(tcur);
        if (dead) return;
        const int l = im.l0 + r8;
        im.R = l < m_loc ? rc_global_row(l, G, rank) : (int64_t)n;
        im.w = im.R < n ? wv.W[(int64_t)im.U * wv.n_pad + im.R] : 0.0;
        im.ok = true;
        flat += TW;
        return;
      }
      base += groups;
      ++tcur;
    }
  };

  Item cur, nxt;
  fetch(cur);
  if (cur.ok) fetch(nxt); else nxt.ok = false;
  unsigned iu = 0, cu = 0;  // units issued / consumed (ring position, mbarrier phase)
  int is = 0;               // issue position: step within cur (or within nxt once in_next)
  bool in_next = false;
  auto issue_one = [&]() {  // warp-uniform bookkeeping; the elected lane issues
    if (!in_next && is >= cur.steps) { in_next = true; is = 0; }
    const Item &im = in_next ? nxt : cur;
    if (!im.ok || is >= im.steps) return;  // (never more than one item ahead)
    if (elect_one()) {
      const unsigned st = iu % kRectStages;
      void *bar = &full[st];
      unsigned char *stg = mystage + (size_t)st * STAGE;
      const int c0 = im.c_base + is * kPanel;
      mbar_arrive_expect_tx(bar, (unsigned)WTILE + 768u);
      tma_load_2d(stg, &tmapA, c0, im.l0, bar);
      if (HALVES == 2) tma_load_2d(stg + 1024, &tmapA, c0 + BOXC, im.l0, bar);
      bulk_g2s(stg + WTILE, im.pay + is * kPanel, 256u, bar);
      bulk_g2s(stg + WTILE + 256, im.pay + kBig + is * kPanel, 256u, bar);
      bulk_g2s(stg + WTILE + 512, im.pay + 2 * kBig + is * kPanel, 256u, bar);
    }
    __syncwarp();
    ++iu;
    ++is;
  };
  if (cur.ok)
    for (int k = 0; k < kRectStages - 1; ++k) issue_one();
  while (cur.ok) {
    double w = cur.w;
    for (int s = 0; s < cur.steps; ++s) {
      const unsigned st = cu % kRectStages;
      mbar_wait(&full[st], (cu / kRectStages) & 1u, status);
      unsigned char *stg = mystage + (size_t)st * STAGE;
      unsigned char *rowp = stg + my_off;
      const double *cf = reinterpret_cast<const double *>(stg + WTILE) + g * 8;
      int4 raw[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) raw[i] = *reinterpret_cast<const int4 *>(rowp + (((cb + i) ^ r8) << 4));
      T *x = reinterpret_cast<T *>(raw);
      double lo[8], pre[8];
      double run = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 a = *reinterpret_cast<const double2 *>(cf + j);
        lo[j] = (double)x[j];
        pre[j] = run;
        run = fma(lo[j], a.x, run);
        lo[j + 1] = (double)x[j + 1];
        pre[j + 1] = run;
        run = fma(lo[j + 1], a.y, run);
      }
      double incl = run;
      double y = __shfl_up_sync(0xffffffffu, incl, 8);
      if (g >= 1) incl += y;
      y = __shfl_up_sync(0xffffffffu, incl, 16);
      if (g >= 2) incl += y;
      const double total = __shfl_sync(0xffffffffu, incl, 24 + r8);
      const double wseg = w - (incl - run);
      w -= total;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 c = *reinterpret_cast<const double2 *>(cf + kPanel + j);
        const double2 sg = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + j);
        x[j] = (T)fma(c.x, lo[j], sg.x * (wseg - pre[j]));
        x[j + 1] = (T)fma(c.y, lo[j + 1], sg.y * (wseg - pre[j + 1]));
      }
#pragma unroll
      for (int i = 0; i < CPT; ++i) *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        const int c0 = cur.c_base + s * kPanel;
        tma_store_2d(&tmapA, c0, cur.l0, stg);
        if (HALVES == 2) tma_store_2d(&tmapA, c0 + BOXC, cur.l0, stg + 1024);
        bulk_commit();
        bulk_wait_read<1>();  // the store of the previous unit has left the stage the next load goes into
      }
      __syncwarp();
      ++cu;
      issue_one();
    }
    if (g == 0 && cur.R < n) wv.W[(int64_t)cur.U * wv.n_pad + cur.R] = w;
    // next item: the issue cursor is already inside it (or waiting at its door)
    if (!in_next) is = 0;
    in_next = false;
    cur = nxt;
    if (cur.ok) fetch(nxt); else nxt.ok = false;
    // (the cursor may have been held at the item boundary: top the ring up again)
    while (cur.ok && iu < cu + (unsigned)(kRectStages - 1)) {
      const unsigned before = iu;
      issue_one();
      if (iu == before) break;
    }
  }
  // (a dead peer: loads already issued still have to land before the CTA may exit)
  for (; cu < iu; ++cu) mbar_wait(&full[cu % kRectStages], (cu / kRectStages) & 1u, status);
  if (elect_one()) bulk_wait_all<0>();
}

}  // namespace chub
