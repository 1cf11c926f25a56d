// dataflow.cuh -- persistent single-launch dataflow kernel for one large factor.
//
// Same arithmetic as large.cuh (GGMS / blocked forward-substitution form of the Seeger update,
// SURVEY appendix A.2), organised so that the lower triangle is streamed through the SMs exactly
// once with the sequential chain overlapped:
//
//   * CTA 0 is the CHAIN.  It never writes L.  For block row r (32 rows) it needs
//         w_r = wt_r - sum_{j=r-d}^{r-1} L[r, j] p_j ,      p_r = inv(L[r, r]) w_r
//     where wt_r ("w-tilde") is the forward-substitution residual of those rows after all block
//     columns j < r-d, handed over by the workers through global memory, and the last d block
//     columns (the BAND) are applied by the chain itself from tiles it prefetched with bulk async
//     copies (TMA engine, cp.async.bulk + mbarrier).  The critical path per 32 columns is two
//     32x32 mat-vecs out of shared memory -- no sqrt/div, no inter-CTA round trip (a spare warp
//     polls wt_{r+1} while the mat-vecs of block column r run).  p_r is published with plain
//     8-byte stores; consumers spin on a NaN sentinel (LL-style: the data is its own flag).
//   * CTAs 1..G-1 are WORKERS.  Rows are dealt to them in blocks of R rows, cyclically, so the
//     shrinking triangle stays balanced.  One thread owns one row for the whole kernel and keeps
//     its residual w_i in a register.  At step s a row processes one 32-column tile: block column
//     s while it is still far left of the diagonal (s < b-d, b = the row's block row), nothing
//     while the chain is inside its band, and block column s-d for s in [b, b+d].  The delay keeps
//     the OLD band entries in memory until the chain has consumed them (the chain publishes p_b
//     only after reading them; step s is gated on p_s).  Each row segment (32 elements) is one
//     bulk async copy into a padded (bank-conflict-free) shared-memory row, prefetched kStages-2
//     steps ahead; it is updated in place
//         L_ik <- c_k L_ik + sigma_k w_i ;  w_i <- w_i - L_ik(old) p_k ;  L_ii <- dn_i
//     and written back with one bulk async store.  Row warps are autonomous (no CTA barriers):
//     a thread loads, updates and stores its own row; mbarriers track the copies.  Warp 7 turns
//     p_s into (c, sigma, dn) ahead of time (prefix sum of p^2; every worker does this
//     redundantly, off the critical path) and hands them over through a full/empty mbarrier ring.
//
// Deadlock freedom: the chain at block row r waits only for workers' step r-d-1; a worker's step
// s waits only for p_s; loads wait for nothing.  All CTAs are co-resident (cooperative launch,
// one CTA per SM).  Every global spin has a watchdog that raises CHUB_STATUS_INTERNAL instead of
// hanging.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; cuTensorMapEncodeTiled is fetched at run time)

#include <algorithm>
#include <cstdlib>

#include "large.cuh"

namespace chub {

constexpr int kDfThreads = 256;
constexpr int kDfRowWarps = 7;  // warps 0..6 own rows, warp 7 derives the coefficients
#ifndef CHUB_DF_D
#define CHUB_DF_D 5
#endif
#ifndef CHUB_DF_SLOTS
#define CHUB_DF_SLOTS 4
#endif
#ifndef CHUB_DF_RING
#define CHUB_DF_RING 16
#endif
constexpr int kDfD = CHUB_DF_D;         // v3 (row-major): band depth in block columns (look-ahead of the chain)
constexpr int kDfStages = 6;            // worker tile ring
constexpr int kDfSlots = CHUB_DF_SLOTS;  // v3: chain prefetch ring (block columns)
constexpr int kDfCoefRing = CHUB_DF_RING;  // v3: coefficient ring
#ifndef CHUB_DF_PF
#define CHUB_DF_PF 6
#endif
constexpr int kDfPf = CHUB_DF_PF;       // v3: chain L2 prefetch distance beyond the ring (block columns)
constexpr int kV2D = 4;         // v2 (column-major) keeps its own, shallower configuration
constexpr int kV2Slots = 4;
constexpr int kV2CoefRing = 8;
constexpr unsigned long long kSentinelBits = 0xFFFFFFFFFFFFFFFFull;
constexpr long long kWatchdogCycles = 4000000000LL;  // ~2 s

struct DfParams {
  int n, T, R, r_shift, Wk, rows_max, nw;  // nw = row warps in use
  int use3d = 0;                           // rm5 workers: the 3-D tensor map (64-column tiles in one TMA op) is valid
  int64_t ld;
  LargeWs ws;
  int32_t *status;
  long long *prof;  // optional [gridDim.x][16] cycle counters (nullptr = off)
  int prof_late = 0;  // CHUB_PROFILE=2: the chain's counters cover the last 3/8 of the block rows only
};

// ------------------------------------------------------------------ primitives
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Wait with a watchdog: a copy that never lands must not hang the GPU.
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity, int32_t *status) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0u) {
      if (clock64() - t0 > 4000000000LL) {
        atomicMax(status, CHUB_STATUS_INTERNAL);
        return;
      }
    }
  }
}
// global -> shared bulk async copy (TMA engine), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk async store, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// all state spaces: generic-proxy accesses (e.g. an acquire of a flag a peer GPU raised) before async-proxy
// copies that read global memory
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// 2-D tensor-map TMA: one instruction moves a whole box (c0 = column, c1 = row of its corner)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, void *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// tensor-map TMA prefetch of a box into L2 (no shared-memory destination, nothing to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, void *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// exactly one lane of a converged warp (the compiler then keeps TMA operands on the uniform path)
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const void *p) {
  unsigned long long x;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void st_relaxed_f64(void *p, double x) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;\n" ::"l"(p), "d"(x) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const void *p) {
  int x;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(x) : "l"(p) : "memory");
  return x;
}

// Spin until *p is not the sentinel.  On watchdog expiry (or if another CTA already gave up)
// flag CHUB_STATUS_INTERNAL and return 0.0 so that the kernel drains instead of hanging.
__device__ __forceinline__ double df_wait_value(const double *p, int32_t *status) {
  unsigned long long bits = ld_relaxed_u64(p);
  if (bits != kSentinelBits) return __longlong_as_double((long long)bits);
  const long long t0 = clock64();
  unsigned spins = 0;
  while (true) {
    bits = ld_relaxed_u64(p);
    if (bits != kSentinelBits) return __longlong_as_double((long long)bits);
    if ((++spins & 255u) == 0u) {
      if (ld_relaxed_s32(status) == CHUB_STATUS_INTERNAL) return 0.0;
      if (clock64() - t0 > kWatchdogCycles) {
        atomicMax(status, CHUB_STATUS_INTERNAL);
        return 0.0;
      }
    }
  }
}

// Poll the value at `cur` with the poll of `nxt` (the next block row's entry, wanted right after) in flight
// at the same time: every round issues both loads back to back, so a consumer that takes one block row
// per round trip keeps up with a producer twice as fast.  b_cur / b_nxt carry polls that are already back.
__device__ __forceinline__ double df_wait_value2(const double *cur, const double *nxt, unsigned long long &b_cur,
                                                 unsigned long long &b_nxt, int32_t *status) {
  if (b_cur == kSentinelBits) {
    const long long t0 = clock64();
    unsigned spins = 0;
    while (true) {
      const unsigned long long x = ld_relaxed_u64(cur);
      const unsigned long long y = nxt ? ld_relaxed_u64(nxt) : kSentinelBits;
      b_nxt = y;
      if (x != kSentinelBits) { b_cur = x; break; }
      if ((++spins & 255u) == 0u) {
        if (ld_relaxed_s32(status) == CHUB_STATUS_INTERNAL) { b_cur = 0; break; }
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(status, CHUB_STATUS_INTERNAL); b_cur = 0; break; }
      }
    }
  }
  return __longlong_as_double((long long)b_cur);
}

__device__ __forceinline__ double df_sanitize(double x) {
  // never let user data alias the sentinel
  return (unsigned long long)__double_as_longlong(x) == kSentinelBits ? __longlong_as_double(0x7FF8000000000000LL) : x;
}

// Which block column does a row of block row b process at step s?  (-1: none)
template <int D>
__device__ __forceinline__ int df_panel_at(int b, int s) {
  if (s < b - D) return s;                          // far left of the diagonal
  if (s >= b && s <= b + D && s >= D) return s - D;  // band, delayed by d steps
  return -1;
}

template <typename T> struct DfGeom {
  static constexpr int EPC = 16 / sizeof(T);          // elements per 16 bytes
  static constexpr int STRIDE = kPanel + EPC;         // padded row of a row-major tile (elements)
  static constexpr int CPR = kPanel / EPC;            // 16-byte chunks per 32-element row segment
};

__device__ __forceinline__ long long df_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// event trace (CHUB_PROFILE): trace[ev * T + block_row] = %globaltimer (ns)
// (PROF is a compile-time flag in scope: the production kernel carries none of this)
#define DF_TRACE(ev, r) do { if (PROF && P.prof) P.prof[256 * 16 + (size_t)(ev) * P.T + (r)] = df_gtime(); } while (0)
#define DF_PROF_T() ((PROF && P.prof) ? clock64() : 0LL)
#define DF_PROF_ADD(slot, t_begin) do { if (PROF && P.prof) pacc[slot] += clock64() - (t_begin); } while (0)

// ------------------------------------------------------------------ prepare
template <typename T, int LAYOUT>
__global__ void df_prepare_kernel(const T *L, int n, int64_t ld, const T *v, LargeWs ws, int flags,
                                  int32_t *status, int band_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pad = (n + 31) & ~31;
  if (i >= n_pad) return;
  const double sentinel = __longlong_as_double((long long)kSentinelBits);
  ws.p[i] = sentinel;
  if (i < n) {
    const double vi = df_sanitize((double)v[i]);
    const T dgi = L[idx<LAYOUT>(i, i, ld)];
    ws.v0[i] = vi;
    ws.dg[i] = (double)dgi;
    ws.w[i] = (i >> 5) <= band_depth ? vi : sentinel;  // rows whose band starts at column 0
    if ((flags & CHUB_FLAG_CHECK_DIAG) && dgi == T(0)) atomicMax(status, CHUB_STATUS_ZERO_DIAG);
  } else {
    ws.v0[i] = 0.0;
    ws.dg[i] = 1.0;
    ws.w[i] = 0.0;
  }
  if (i == 0) { ws.scal[0] = 1.0; ws.scal[1] = 1.0; ws.scal[2] = 0.0; ws.scal[3] = 0.0; }
}

// Fused pre-pass: block b of 128 threads prepares entries [128 b, 128 b + 128) (= the four diagonal
// blocks it inverts).  check_diag failures are only known after the whole grid ran, so the inverses
// are computed unconditionally (they only read L) -- the persistent kernel checks the status word.
template <typename T, int LAYOUT>
// w_in != nullptr (blocked driver, diagonal block J > 0): the vector entering this block is the residual
// left in ws.w by the panels to its left, and (t, S) carry over from the previous block (ws.scal[4..5]).
__global__ void __launch_bounds__(128) df_prepass_kernel(const T *L, int n, int64_t ld, const T *v, LargeWs ws,
                                                         int flags, int32_t *status, int band_depth,
                                                         bool with_m1, const double *w_in) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int n_pad = (n + 31) & ~31;
  if (i < n_pad) {
    const double sentinel = __longlong_as_double((long long)kSentinelBits);
    ws.p[i] = sentinel;
    if (i < n) {
      const double vi = df_sanitize(w_in ? w_in[i] : (double)v[i]);
      const T dgi = L[idx<LAYOUT>(i, i, ld)];
      ws.v0[i] = vi;
      ws.dg[i] = (double)dgi;
      ws.w[i] = (i >> 5) <= band_depth ? vi : sentinel;
      if ((flags & CHUB_FLAG_CHECK_DIAG) && dgi == T(0)) atomicMax(status, CHUB_STATUS_ZERO_DIAG);
    } else {
      ws.v0[i] = 0.0;
      ws.dg[i] = 1.0;
      ws.w[i] = 0.0;
    }
  }
  if (i == 0) {
    ws.scal[0] = w_in ? ws.scal[4] : 1.0;
    ws.scal[1] = w_in ? ws.scal[5] : 1.0;
    ws.scal[2] = 0.0; ws.scal[3] = 0.0;
  }
  invdiag_block<T, LAYOUT>(L, n, ld, ws, blockIdx.x * 4 + (threadIdx.x >> 5), with_m1);
}

// ------------------------------------------------------------------ chain CTA
// Shared-memory tile of the band: row-major source -> [32 rows][STRIDE] (lane = row reads 16-byte
// chunks conflict-free thanks to the padding); column-major source -> [32 cols][32 rows].
template <typename T, int LAYOUT>
__device__ void df_chain(const T *L, const DfParams &P, unsigned char *smem) {
  constexpr bool PROF = true;
  constexpr int EPC = DfGeom<T>::EPC, CPR = DfGeom<T>::CPR, STRIDE = DfGeom<T>::STRIDE;
  constexpr int TILE_ELEMS = LAYOUT == CHUB_ROW_MAJOR ? kPanel * STRIDE : kPanel * kPanel;
  constexpr int LOADERS = kDfThreads - kV2D * 32;  // warps d..7
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = P.n, T_ = P.T;
  const int64_t ld = P.ld;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [kV2Slots]
  double *Xs = reinterpret_cast<double *>(smem + 64);                       // [kV2Slots][kXBlock]
  double *acc = Xs + kV2Slots * kXBlock;                                    // [kV2D+1][32]
  double *pbuf = acc + (kV2D + 1) * kPanel;                                 // [2][32]
  double *wvec = pbuf + 2 * kPanel;                                         // [32]
  double *wtb = wvec + kPanel;                                              // [2][32] polled wt
  T *tiles = reinterpret_cast<T *>(wtb + 2 * kPanel);                       // [kV2Slots][kV2D][TILE_ELEMS]

  if (tid == 0) {
    for (int i = 0; i < kV2Slots; ++i) mbar_init(&full[i], LOADERS);
    fence_mbar_init();
  }
  for (int e = tid; e < (kV2D + 1) * kPanel; e += kDfThreads) acc[e] = 0.0;
  __syncthreads();

  // Loads for block column j: tiles (j+1 .. j+d, j) and X_{j+1}, into slot j % kV2Slots.
  // Called by the LOADERS threads (t = 0 .. LOADERS-1); each arrives exactly once.
  auto issue_slot = [&](int j, int t) {
    if (j < 0 || j >= T_) return;
    const int slot = j % kV2Slots;
    void *bar = &full[slot];
    unsigned bytes = 0;
    const int q = t >> 5, rr = t & 31;  // tile q (block row j+1+q), row / column rr of it
    const void *src = nullptr;
    void *dst = nullptr;
    if (q < kV2D) {
      const int br = j + 1 + q;
      T *tb = tiles + ((size_t)slot * kV2D + q) * TILE_ELEMS;
      if (br < T_) {
        if (LAYOUT == CHUB_ROW_MAJOR) {
          const int64_t row = (int64_t)br * kPanel + rr;
          if (row < n) {
            src = L + row * ld + (int64_t)j * kPanel;
            dst = tb + rr * STRIDE;
            bytes = kPanel * sizeof(T);
          }
        } else {
          const int64_t row0 = (int64_t)br * kPanel;
          const int64_t col = (int64_t)j * kPanel + rr;
          int rows_in = (int)min((int64_t)kPanel, (int64_t)n - row0);
          rows_in = (rows_in + EPC - 1) / EPC * EPC;
          if (rows_in > 0) {
            src = L + row0 + col * ld;
            dst = tb + rr * kPanel;
            bytes = rows_in * sizeof(T);
          }
        }
      }
    }
    const bool do_x = (t == 0) && (j + 1 < T_);
    const unsigned xbytes = do_x ? kXBlock * sizeof(double) : 0u;
    if (bytes + xbytes) mbar_arrive_expect_tx(bar, bytes + xbytes);
    else mbar_arrive(bar);
    if (bytes) bulk_g2s(dst, src, bytes, bar);
    if (do_x) bulk_g2s(Xs + slot * kXBlock, P.ws.X + (int64_t)(j + 1) * kXBlock, xbytes, bar);
  };

  // prologue: prefetch slots 0 .. kV2Slots-2; p_0 = X_0 w_0 (w_0 = v[0:32])
  if (warp >= kV2D)
    for (int j = 0; j < kV2Slots - 1; ++j) issue_slot(j, tid - kV2D * 32);
  if (warp == 0) {
    wvec[lane] = df_wait_value(P.ws.w + lane, P.status);
    __syncwarp();
    const double *X0 = P.ws.X + lane * kXStride;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
    for (int k = 0; k < kPanel; k += 2) {
      a0 += X0[k] * wvec[k];
      a1 += X0[k + 1] * wvec[k + 1];
    }
    const double p0 = a0 + a1;
    pbuf[lane] = p0;
    st_relaxed_f64(P.ws.p + lane, p0);
  }
  __syncthreads();

  const long long t_chain0 = DF_PROF_T();
  for (int j = 0; j < T_; ++j) {
    // ---- phase A: acc_{j+1+q} -= L[j+1+q, j] p_j  (warp q < d); warps d..7 prefetch; warp d
    //      also polls wt_{j+1} so that phase B finds it in shared memory
    long long tp = DF_PROF_T();
    const double *pj = pbuf + (j & 1) * kPanel;
    if (warp < kV2D) {
      const int br = j + 1 + warp;
      if (br < T_) {
        mbar_wait(&full[j % kV2Slots], (unsigned)((j / kV2Slots) & 1), P.status);
        if (tid == 0) DF_PROF_ADD(1, tp);
        const T *tb = tiles + ((size_t)(j % kV2Slots) * kV2D + warp) * TILE_ELEMS;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if ((int64_t)br * kPanel + lane < n) {
          if (LAYOUT == CHUB_ROW_MAJOR) {
            const T *rowp = tb + lane * STRIDE;
#pragma unroll
            for (int ch = 0; ch < CPR; ++ch) {
              if (EPC == 2) {
                const double2 x = *reinterpret_cast<const double2 *>(rowp + ch * 2);
                if (ch & 1) { a2 += x.x * pj[ch * 2]; a3 += x.y * pj[ch * 2 + 1]; }
                else        { a0 += x.x * pj[ch * 2]; a1 += x.y * pj[ch * 2 + 1]; }
              } else {
                const float4 x = *reinterpret_cast<const float4 *>(rowp + ch * 4);
                a0 += (double)x.x * pj[ch * 4];
                a1 += (double)x.y * pj[ch * 4 + 1];
                a2 += (double)x.z * pj[ch * 4 + 2];
                a3 += (double)x.w * pj[ch * 4 + 3];
              }
            }
          } else {
#pragma unroll 8
            for (int k = 0; k < kPanel; k += 4) {
              a0 += (double)tb[k * kPanel + lane] * pj[k];
              a1 += (double)tb[(k + 1) * kPanel + lane] * pj[k + 1];
              a2 += (double)tb[(k + 2) * kPanel + lane] * pj[k + 2];
              a3 += (double)tb[(k + 3) * kPanel + lane] * pj[k + 3];
            }
          }
        }
        acc[(br % (kV2D + 1)) * kPanel + lane] -= (a0 + a1) + (a2 + a3);
      }
      if (tid == 0) DF_PROF_ADD(2, tp);
    } else {
      // refill the slot that phase A of block column j-1 released at the last barrier
      issue_slot(j + kV2Slots - 1, tid - kV2D * 32);
      if (warp == kV2D && j + 1 < T_) {
        long long tq = DF_PROF_T();
        wtb[((j + 1) & 1) * kPanel + lane] =
            df_wait_value(P.ws.w + (int64_t)(j + 1) * kPanel + lane, P.status);
        if (lane == 0) DF_PROF_ADD(4, tq);
      }
    }
    tp = DF_PROF_T();
    __syncthreads();
    // ---- phase B: w_{j+1} = wt_{j+1} + acc ; p_{j+1} = X_{j+1} w_{j+1}
    if (warp == 0 && j + 1 < T_) {
      const int r = j + 1;
      double *a = acc + (r % (kV2D + 1)) * kPanel;
      wvec[lane] = wtb[(r & 1) * kPanel + lane] + a[lane];
      a[lane] = 0.0;
      __syncwarp();
      if (tid == 0) DF_PROF_ADD(3, tp);
      tp = DF_PROF_T();
      const double *X = Xs + (j % kV2Slots) * kXBlock + lane * kXStride;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int k = 0; k < kPanel; k += 4) {
        const double2 x = *reinterpret_cast<const double2 *>(X + k);
        const double2 y = *reinterpret_cast<const double2 *>(X + k + 2);
        a0 += x.x * wvec[k];
        a1 += x.y * wvec[k + 1];
        a2 += y.x * wvec[k + 2];
        a3 += y.y * wvec[k + 3];
      }
      const double pr = (a0 + a1) + (a2 + a3);
      pbuf[(r & 1) * kPanel + lane] = pr;
      st_relaxed_f64(P.ws.p + (int64_t)r * kPanel + lane, pr);  // publish: data is its own flag
      if (tid == 0) DF_PROF_ADD(5, tp);
    }
    __syncthreads();
  }
  if (tid == 0) DF_PROF_ADD(0, t_chain0);
  if (P.prof && (tid == 0 || tid == kV2D * 32)) {
    for (int i = 0; i < 8; ++i)
      if (pacc[i]) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_chain_smem(int layout) {
  const size_t tile_elems = layout == CHUB_ROW_MAJOR ? (size_t)kPanel * DfGeom<T>::STRIDE : (size_t)kPanel * kPanel;
  return 64 + (size_t)(kV2Slots * kXBlock + (kV2D + 1) * kPanel + 5 * kPanel) * sizeof(double) +
         (size_t)kV2Slots * kV2D * tile_elems * sizeof(T);
}

// ------------------------------------------------------------------ worker CTAs
template <typename T, int LAYOUT, bool UPDATE>
__device__ void df_worker(T *L, T *v, const DfParams &P, unsigned char *smem) {
  constexpr bool PROF = true;
  constexpr int EPC = DfGeom<T>::EPC, CPR = DfGeom<T>::CPR, STRIDE = DfGeom<T>::STRIDE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wkr = blockIdx.x - 1;
  const int n = P.n, T_ = P.T, R = P.R, rsh = P.r_shift, Wk = P.Wk, nw = P.nw;
  const int64_t ld = P.ld;
  const int nsteps = T_ + kV2D;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  // shared memory carve-up
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [kDfStages][kDfRowWarps]
  unsigned long long *cfull = full + kDfStages * kDfRowWarps;               // [kV2CoefRing]
  unsigned long long *cempty = cfull + kV2CoefRing;                         // [kV2CoefRing]
  double *coef = reinterpret_cast<double *>(smem + 512);                    // [kV2CoefRing][4][32]
  T *stages = reinterpret_cast<T *>(coef + kV2CoefRing * 4 * kPanel);
  // row-major: stage = [nw*32 rows][STRIDE];  column-major: stage = [nw warps][32 cols][32 rows]
  const size_t stage_elems = (size_t)nw * 32 * (LAYOUT == CHUB_ROW_MAJOR ? STRIDE : kPanel);

  if (tid == 0) {
    for (int i = 0; i < kDfStages * kDfRowWarps; ++i) mbar_init(&full[i], 32);
    for (int i = 0; i < kV2CoefRing; ++i) { mbar_init(&cfull[i], 32); mbar_init(&cempty[i], nw); }
    fence_mbar_init();
  }
  __syncthreads();

  auto row_of = [&](int lr) -> int64_t {
    return (((int64_t)wkr + (int64_t)(lr >> rsh) * Wk) << rsh) + (lr & (R - 1));
  };

  if (warp == 7) {
    // ---------------- coefficient warp: p_s -> (p, c, sigma, dn) of block column s
    double t_run = 1.0, S_run = 1.0;
    for (int pan = 0; pan < T_; ++pan) {
      const int slot = pan % kV2CoefRing;
      if (pan >= kV2CoefRing) mbar_wait(&cempty[slot], (unsigned)((pan / kV2CoefRing - 1) & 1), P.status);
      const int k = pan * kPanel + lane;
      long long tq = DF_PROF_T();
      const double pk = df_wait_value(P.ws.p + k, P.status);
      if (lane == 0) DF_PROF_ADD(6, tq);
      double *cf = coef + (size_t)slot * 4 * kPanel;
      cf[lane] = pk;
      if (UPDATE) {
        const double dg = P.ws.dg[k];
        double incl = pk * pk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        const double tk = t_run + excl, tk1 = t_run + incl;
        const unsigned negmask = __ballot_sync(0xffffffffu, dg < 0.0);
        const double Sk = (__popc(negmask & ((1u << lane) - 1u)) & 1) ? -S_run : S_run;
        const double sgn = dg < 0.0 ? -1.0 : 1.0;
        const double rinv = 1.0 / sqrt(tk * tk1);
        const double ck = sgn * tk * rinv, sigk = sgn * pk * rinv;
        cf[kPanel + lane] = ck;
        cf[2 * kPanel + lane] = sigk;
        cf[3 * kPanel + lane] = fabs(dg) * tk1 * rinv;
        if (wkr == pan % Wk && k < n) {  // one worker per panel leaves rotg's z_k in v[k]
          const double nuk = Sk * dg * pk / sqrt(tk);
          const double sk = Sk * sgn * pk / sqrt(tk1);
          v[k] = (T)rotg_z<double>(dg, nuk, ck, sk);
        }
        t_run = __shfl_sync(0xffffffffu, tk1, 31);
        if (__popc(negmask) & 1) S_run = -S_run;
      }
      mbar_arrive(&cfull[slot]);  // release: the 32 lanes' shared-memory writes become visible
      if (lane == 0) DF_PROF_ADD(7, tq);
    }
    if (P.prof && lane == 0) {
      P.prof[(size_t)blockIdx.x * 16 + 6] = pacc[6];
      P.prof[(size_t)blockIdx.x * 16 + 7] = pacc[7];
    }
    return;
  }
  if (warp >= nw) return;

  // ---------------- row warps
  const int lr = warp * 32 + lane;          // local row
  const bool prof_me = (warp == nw - 1) && lane == 0;  // the warp that holds the bottom rows
  const int64_t row = row_of(lr);
  const bool has_row = lr < P.rows_max && row < n;
  const int b = has_row ? (int)(row >> 5) : -1000000;
  double w = has_row ? P.ws.v0[row] : 0.0;
  // column-major: lane k moves column k of the two 16-row blocks (h = 0, 1) of this warp
  int64_t cm_row0[2] = {0, 0};
  int cm_b[2] = {-1000000, -1000000}, cm_rows[2] = {0, 0};
  if (LAYOUT == CHUB_COL_MAJOR) {
    for (int h = 0; h < 2; ++h) {
      const int lr0 = warp * 32 + 16 * h;
      const int64_t r0 = row_of(lr0);
      if (lr0 < P.rows_max && r0 < n) {
        cm_row0[h] = r0;
        cm_b[h] = (int)(r0 >> 5);
        cm_rows[h] = (int)min((int64_t)16, (int64_t)n - r0);
      }
    }
  }

  auto issue_load = [&](int s) {
    if (s >= nsteps) return;
    void *bar = &full[(s % kDfStages) * kDfRowWarps + warp];
    T *st = stages + (size_t)(s % kDfStages) * stage_elems;
    if (LAYOUT == CHUB_ROW_MAJOR) {
      const int pan = has_row ? df_panel_at<kV2D>(b, s) : -1;
      if (pan >= 0) {
        const int64_t c0 = (int64_t)pan * kPanel;
        int cnt = (int)min((int64_t)kPanel, min((int64_t)n - c0, row - c0 + 1));
        cnt = (cnt + EPC - 1) / EPC * EPC;
        const unsigned bytes = cnt * sizeof(T);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(st + (size_t)lr * STRIDE, L + row * ld + c0, bytes, bar);
      } else {
        mbar_arrive(bar);
      }
    } else {
      unsigned total = 0;
      const void *src[2] = {nullptr, nullptr};
      void *dst[2] = {nullptr, nullptr};
      unsigned by[2] = {0, 0};
      T *wt = st + (size_t)warp * kPanel * kPanel;
      for (int h = 0; h < 2; ++h) {
        const int pan = df_panel_at<kV2D>(cm_b[h], s);
        if (pan < 0) continue;
        const int64_t col = (int64_t)pan * kPanel + lane;
        if (col >= n || col > cm_row0[h] + cm_rows[h] - 1) continue;  // outside the lower triangle
        const int rows_r = (cm_rows[h] + EPC - 1) / EPC * EPC;
        src[h] = L + cm_row0[h] + col * ld;
        dst[h] = wt + lane * kPanel + 16 * h;
        by[h] = rows_r * sizeof(T);
        total += by[h];
      }
      if (total) mbar_arrive_expect_tx(bar, total);
      else mbar_arrive(bar);
      for (int h = 0; h < 2; ++h)
        if (by[h]) bulk_g2s(dst[h], src[h], by[h], bar);
    }
  };

  for (int s = 0; s < kDfStages - 2; ++s) issue_load(s);

  const long long t_wk0 = DF_PROF_T();
  for (int s = 0; s < nsteps; ++s) {
    long long tp = DF_PROF_T();
    if (LAYOUT == CHUB_COL_MAJOR) __syncwarp();  // lanes share the warp's tile: all done with step s-1
    bulk_wait_read<1>();  // my store of step s-2 has left shared memory: its stage is free
    issue_load(s + kDfStages - 2);
    if (prof_me) DF_PROF_ADD(2, tp);
    tp = DF_PROF_T();
    mbar_wait(&full[(s % kDfStages) * kDfRowWarps + warp], (unsigned)((s / kDfStages) & 1), P.status);
    if (prof_me) DF_PROF_ADD(1, tp);
    tp = DF_PROF_T();
    if (s < T_) mbar_wait(&cfull[s % kV2CoefRing], (unsigned)((s / kV2CoefRing) & 1), P.status);
    if (prof_me) DF_PROF_ADD(4, tp);
    tp = DF_PROF_T();

    const int pan = has_row ? df_panel_at<kV2D>(b, s) : -1;
    T *st = stages + (size_t)(s % kDfStages) * stage_elems;
    if (pan >= 0) {
      const double *cf = coef + (size_t)(pan % kV2CoefRing) * 4 * kPanel;
      const int64_t c0 = (int64_t)pan * kPanel;
      const bool full_tile = c0 + kPanel <= row;  // whole tile strictly below the diagonal
      if (LAYOUT == CHUB_ROW_MAJOR) {
        T *rowp = st + (size_t)lr * STRIDE;
#pragma unroll 4
        for (int ch = 0; ch < CPR; ++ch) {
          int4 raw = *reinterpret_cast<const int4 *>(rowp + ch * EPC);
          T *x = reinterpret_cast<T *>(&raw);
#pragma unroll
          for (int u = 0; u < EPC; ++u) {
            const int k = ch * EPC + u;
            const int64_t col = c0 + k;
            if (full_tile || col < row) {
              const double lo = (double)x[u];
              if (UPDATE) x[u] = (T)(cf[kPanel + k] * lo + cf[2 * kPanel + k] * w);
              w -= lo * cf[k];
            } else if (UPDATE && col == row) {
              x[u] = (T)cf[3 * kPanel + k];
            }
          }
          if (UPDATE) *reinterpret_cast<int4 *>(rowp + ch * EPC) = raw;
        }
      } else {
        T *colp = st + (size_t)warp * kPanel * kPanel + lane;
#pragma unroll 8
        for (int k = 0; k < kPanel; ++k) {
          T *sm = colp + k * kPanel;
          const int64_t col = c0 + k;
          if (full_tile || col < row) {
            const double lo = (double)*sm;
            if (UPDATE) *sm = (T)(cf[kPanel + k] * lo + cf[2 * kPanel + k] * w);
            w -= lo * cf[k];
          } else if (UPDATE && col == row) {
            *sm = (T)cf[3 * kPanel + k];
          }
        }
      }
      // hand the residual over to the chain after the last far-left block column
      if (s == b - kV2D - 1) st_relaxed_f64(P.ws.w + row, df_sanitize(w));
    }
    if (prof_me) DF_PROF_ADD(3, tp);
    tp = DF_PROF_T();

    fence_proxy_async();  // order my generic-proxy accesses before later bulk (async proxy) copies
    if (UPDATE) {
      if (LAYOUT == CHUB_ROW_MAJOR) {
        if (pan >= 0) {
          const int64_t c0 = (int64_t)pan * kPanel;
          const int cnt = (int)min((int64_t)kPanel, row - c0 + 1);  // never above the diagonal
          const int cb = cnt / EPC * EPC;
          T *rowp = st + (size_t)lr * STRIDE;
          T *g = L + row * ld + c0;
          if (cb) bulk_s2g(g, rowp, cb * sizeof(T));
          for (int u = cb; u < cnt; ++u) g[u] = rowp[u];
        }
      } else {
        __syncwarp();  // lane k stores column k: entries written by the other lanes of this warp
        T *wt = st + (size_t)warp * kPanel * kPanel;
        for (int h = 0; h < 2; ++h) {
          const int panh = df_panel_at<kV2D>(cm_b[h], s);
          if (panh < 0) continue;
          const int64_t col = (int64_t)panh * kPanel + lane;
          const int64_t r_hi = cm_row0[h] + cm_rows[h];
          int64_t r_lo = max(cm_row0[h], col);
          if (col >= n || r_lo >= r_hi) continue;
          T *sm = wt + lane * kPanel + 16 * h - cm_row0[h];  // sm[r] = element (r, col)
          T *g = L + col * ld;                               // g[r]  = element (r, col)
          while (r_lo < r_hi && (r_lo & (EPC - 1))) { g[r_lo] = sm[r_lo]; ++r_lo; }
          const int64_t r_mid = r_lo + (r_hi - r_lo) / EPC * EPC;
          if (r_mid > r_lo) bulk_s2g(g + r_lo, sm + r_lo, (unsigned)((r_mid - r_lo) * sizeof(T)));
          for (int64_t r = r_mid; r < r_hi; ++r) g[r] = sm[r];
        }
      }
      bulk_commit();
    }
    // this warp is done with block column s-d for good
    if (s >= kV2D && s - kV2D < T_) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&cempty[(s - kV2D) % kV2CoefRing]);
    }
    if (prof_me) DF_PROF_ADD(5, tp);
  }
  if (prof_me) DF_PROF_ADD(0, t_wk0);
  bulk_wait_all<0>();
  if (P.prof && prof_me) {
    for (int i = 0; i < 6; ++i) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_worker_smem(int layout, int nw) {
  const size_t stage_elems = (size_t)nw * 32 * (layout == CHUB_ROW_MAJOR ? DfGeom<T>::STRIDE : kPanel);
  return 512 + (size_t)kV2CoefRing * 4 * kPanel * sizeof(double) + (size_t)kDfStages * stage_elems * sizeof(T);
}

// ------------------------------------------------------------------ the kernel
template <typename T, int LAYOUT, bool UPDATE>
__global__ void __launch_bounds__(kDfThreads, 1) df_kernel_v2(T *L, T *v, DfParams P) {
  extern __shared__ __align__(1024) unsigned char df_smem[];
  if (*P.status != 0) return;  // check_diag failed: L and v stay untouched
  if (blockIdx.x == 0) df_chain<T, LAYOUT>(L, P, df_smem);
  else df_worker<T, LAYOUT, UPDATE>(L, v, P, df_smem);
}

// =====================================================================================
// v3 (row-major): tensor-map TMA boxes + 4 threads per row
// =====================================================================================
// A worker's rows come in blocks of 16 (two warps of 8 rows each).  Per step a warp moves its
// [8 rows x 32 cols] tile with one TMA instruction per 128-byte-wide box (SWIZZLE_128B, so that
// "8 lanes = 8 rows, same 16-byte chunk" is bank-conflict free) issued by lane 0.  Lane (g, r8)
// owns columns 8g..8g+7 of row r8: it forms the local prefix sums of L_ik p_k, the four segment
// totals are combined by a two-step shuffle scan, and every entry is updated from
//     L_ik <- c_k L_ik + sigma_k (w_i - sum_{m<k} L_im p_m)
// so the serial chain per row is 8 long instead of 32 and 4x as many warps hide latency.
constexpr int kRmMaxQ = 7;                       // row blocks (16 rows) per worker
constexpr int kRmMaxThreads = 32 * (2 * kRmMaxQ + 2);  // 14 row warps + coefficient warp (+1 spare) = 512

template <typename T> struct RmGeom {
  static constexpr int HALVES = kPanel * sizeof(T) / 128;  // 128-byte boxes per 32-column tile (2 / 1)
  static constexpr int BOXC = kPanel / HALVES;             // box columns (16 / 32)
  static constexpr int EPC = 16 / sizeof(T);               // elements per 16-byte chunk (2 / 4)
  static constexpr int CPT = 8 / EPC;                      // chunks per thread (8 columns): 4 / 2
  static constexpr int WTILE = HALVES * 1024;              // bytes of a warp's tile (8 rows)
  static constexpr int CTILE = HALVES * 4096;              // bytes of a chain tile (32 rows)
};

// ---- chain CTA (threads 0..255 of CTA 0)
// LAYOUT: row-major band tiles arrive as two 128-byte-swizzled boxes of [32 rows][16 columns] (lane = row
// reads 16-byte chunks, conflict free); column-major ones as one plain box [32 columns][32 rows] (lane = row
// reads 8 bytes per column, conflict free).
template <typename T, int LAYOUT, bool PROF>
__device__ void df_chain_rm(const DfParams &P, const CUtensorMap *tmap32, unsigned char *smem) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, EPC = RmGeom<T>::EPC;
  constexpr int CTILE = RmGeom<T>::CTILE;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int T_ = P.T;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  // (Measured and dropped, profiles/r02_chain_v4_experiments.txt: a ring of their own for the inverse diagonal
  // blocks, loaded two block rows ahead, which would make room for d = 6 band tiles -- phase B then waits for
  // X on the critical path: 0.53 ms per update instead of 0.47.)
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [kDfSlots]
  double *acc = reinterpret_cast<double *>(smem + 64);                      // [kDfD+1][32]
  double *pbuf = acc + (kDfD + 1) * kPanel;                                 // [2][32]
  double *wvec = pbuf + 2 * kPanel;                                         // [32]
  double *wtb = wvec + kPanel;                                              // [2][32]
  double *Xs = reinterpret_cast<double *>(smem + 4096);                     // [kDfSlots][kXBlock]
  unsigned char *tiles = smem + 4096 + ((kDfSlots * kXBlock * 8 + 1023) / 1024) * 1024;  // [kDfSlots][kDfD][CTILE]

  if (tid == 0) {
    for (int i = 0; i < kDfSlots; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  for (int e = tid; e < (kDfD + 1) * kPanel; e += 256) acc[e] = 0.0;
  named_sync(1, 256);

  // one thread issues everything block column j needs: the band tiles (j+1..j+d, j) as ONE box of 32 d rows
  // (ten 2-D boxes issued one after the other by one lane cost ~1000 cycles per block row: the barrier of
  // every block row waited for them) and X_{j+1}.  Rows past the end of the matrix are zero-filled.
  // Shared-memory layout of a slot: row-major [half h][tile q][32 rows][128 B swizzled];
  // column-major [32 columns][32 d rows].
  auto issue_slot = [&](int j) {
    if (j < 0 || j >= T_) return;
    const int slot = j % kDfSlots;
    void *bar = &full[slot];
    const bool do_t = j + 1 < T_;
    if (!do_t) { mbar_arrive(bar); return; }
    mbar_arrive_expect_tx(bar, (unsigned)kDfD * CTILE + kXBlock * 8u);
    if (LAYOUT == CHUB_COL_MAJOR)  // (coordinates: row first -- rows are the contiguous dimension)
      tma_load_2d(tiles + (size_t)slot * kDfD * CTILE, tmap32, (j + 1) * kPanel, j * kPanel, bar);
    else
      tma_load_3d(tiles + (size_t)slot * kDfD * CTILE, tmap32, 0, (j + 1) * kPanel, j * HALVES, bar);
    bulk_g2s(Xs + slot * kXBlock, P.ws.X + (int64_t)(j + 1) * kXBlock, kXBlock * 8u, bar);
  };

  if (warp == kDfD + 1 && elect_one())
    for (int j = 0; j < kDfSlots - 1; ++j) issue_slot(j);
  if (warp == 0) {
    wvec[lane] = df_wait_value(P.ws.w + lane, P.status);
    __syncwarp();
    const double *X0 = P.ws.X + lane * kXStride;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
    for (int k = 0; k < kPanel; k += 2) {
      a0 += X0[k] * wvec[k];
      a1 += X0[k + 1] * wvec[k + 1];
    }
    const double p0 = a0 + a1;
    pbuf[lane] = p0;
    st_relaxed_f64(P.ws.p + lane, p0);
  }
  // warp d polls wt one block row ahead without blocking: the load for wt_{j+2} is in flight
  // while block column j is processed
  unsigned long long wt_bits = kSentinelBits;
  if (warp == kDfD && T_ > 1) wt_bits = ld_relaxed_u64(P.ws.w + kPanel + lane);
  named_sync(1, 256);

  long long t_chain0 = DF_PROF_T();
  for (int j = 0; j < T_; ++j) {
    if (PROF && P.prof && P.prof_late && j == T_ * 5 / 8) {
      for (int i = 0; i < 8; ++i) pacc[i] = 0;
      t_chain0 = clock64();
    }
    long long tp = DF_PROF_T();
    const double *pj = pbuf + (j & 1) * kPanel;
    if (warp < kDfD) {
      const int br = j + 1 + warp;
      if (br < T_) {
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        if (tid == 0) DF_PROF_ADD(1, tp);
        const unsigned char *sb = tiles + (size_t)(j % kDfSlots) * kDfD * CTILE;  // this block column's slot
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (LAYOUT == CHUB_COL_MAJOR) {
          // element (tile q = warp, row = lane, column k) at [k][32 q + lane]
          const T *tc = reinterpret_cast<const T *>(sb) + warp * kPanel + lane;
          constexpr int CS = kDfD * kPanel;  // column stride (elements)
#pragma unroll
          for (int k = 0; k < kPanel; k += 4) {
            const double2 pp = *reinterpret_cast<const double2 *>(pj + k);
            const double2 pq = *reinterpret_cast<const double2 *>(pj + k + 2);
            a0 += (double)tc[k * CS] * pp.x;
            a1 += (double)tc[(k + 1) * CS] * pp.y;
            a2 += (double)tc[(k + 2) * CS] * pq.x;
            a3 += (double)tc[(k + 3) * CS] * pq.y;
          }
        } else {
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const unsigned char *rowp = sb + h * (kDfD * 4096) + warp * 4096 + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int4 raw = *reinterpret_cast<const int4 *>(rowp + ((c ^ (lane & 7)) << 4));
            const T *x = reinterpret_cast<const T *>(&raw);
            const int k0 = h * BOXC + c * EPC;
            if (EPC == 2) {
              const double2 pp = *reinterpret_cast<const double2 *>(pj + k0);
              if (c & 1) { a2 += (double)x[0] * pp.x; a3 += (double)x[1] * pp.y; }
              else       { a0 += (double)x[0] * pp.x; a1 += (double)x[1] * pp.y; }
            } else {
              const double2 pp = *reinterpret_cast<const double2 *>(pj + k0);
              const double2 pq = *reinterpret_cast<const double2 *>(pj + k0 + 2);
              a0 += (double)x[0] * pp.x;
              a1 += (double)x[1] * pp.y;
              a2 += (double)x[2 % EPC] * pq.x;
              a3 += (double)x[3 % EPC] * pq.y;
            }
          }
        }
        }
        acc[(br % (kDfD + 1)) * kPanel + lane] -= (a0 + a1) + (a2 + a3);
      }
      if (tid == 0) DF_PROF_ADD(2, tp);
    } else if (warp == kDfD) {
      if (j + 1 < T_) {
        long long tq = DF_PROF_T();
#ifdef CHUB_EXP_CHAIN_FREE  // timing experiment (results are garbage): the chain never waits for the workers
        const double wt = wt_bits != kSentinelBits ? __longlong_as_double((long long)wt_bits) : 0.0;
#else
        const double wt = wt_bits != kSentinelBits
                              ? __longlong_as_double((long long)wt_bits)
                              : df_wait_value(P.ws.w + (int64_t)(j + 1) * kPanel + lane, P.status);
#endif
        wtb[((j + 1) & 1) * kPanel + lane] = wt;
        if (j + 2 < T_) wt_bits = ld_relaxed_u64(P.ws.w + (int64_t)(j + 2) * kPanel + lane);
        __syncwarp();
        if (lane == 0) { DF_PROF_ADD(4, tq); DF_TRACE(1, j + 1); }
      }
    } else if (warp == kDfD + 1) {
      if (elect_one()) issue_slot(j + kDfSlots - 1);
    } else if (kDfPf > 0) {
      // the spare warp: the ring's lead (kDfSlots-1 block columns) is shorter than the DRAM latency
      // while every SM streams, so the band tiles of block column j + kDfSlots-1 + kDfPf go to L2 now
      // (issued by the TMA-issuing warp itself it costs 12-50 us per update: measured twice)
      const int jp = j + kDfSlots - 1 + kDfPf;
      if (jp + 1 < T_ && elect_one()) {
        if (LAYOUT == CHUB_COL_MAJOR) tma_prefetch_2d(tmap32, (jp + 1) * kPanel, jp * kPanel);
        else tma_prefetch_3d(tmap32, 0, (jp + 1) * kPanel, jp * HALVES);
      }
    }
    tp = DF_PROF_T();
    named_sync(1, 256);
    if (warp == 0 && j + 1 < T_) {
      const int r = j + 1;
      double *a = acc + (r % (kDfD + 1)) * kPanel;
      wvec[lane] = wtb[(r & 1) * kPanel + lane] + a[lane];
      a[lane] = 0.0;
      __syncwarp();
      if (tid == 0) DF_PROF_ADD(3, tp);
      tp = DF_PROF_T();
      const double *X = Xs + (j % kDfSlots) * kXBlock + lane * kXStride;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int k = 0; k < kPanel; k += 4) {
        const double2 x = *reinterpret_cast<const double2 *>(X + k);
        const double2 y = *reinterpret_cast<const double2 *>(X + k + 2);
        const double2 u = *reinterpret_cast<const double2 *>(wvec + k);
        const double2 z = *reinterpret_cast<const double2 *>(wvec + k + 2);
        a0 += x.x * u.x;
        a1 += x.y * u.y;
        a2 += y.x * z.x;
        a3 += y.y * z.y;
      }
      const double pr = (a0 + a1) + (a2 + a3);
      pbuf[(r & 1) * kPanel + lane] = pr;
      st_relaxed_f64(P.ws.p + (int64_t)r * kPanel + lane, pr);
      if (tid == 0) { DF_PROF_ADD(5, tp); DF_TRACE(2, r); }
    }
    named_sync(1, 256);
  }
  if (tid == 0) DF_PROF_ADD(0, t_chain0);
  if (P.prof && (tid == 0 || tid == kDfD * 32)) {
    for (int i = 0; i < 8; ++i)
      if (pacc[i]) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_chain_rm_smem() {
  return 4096 + ((size_t)(kDfSlots * kXBlock * 8 + 1023) / 1024) * 1024 +
         (size_t)kDfSlots * kDfD * RmGeom<T>::CTILE;
}

// ---- coefficient warp of a worker CTA (shared by the row-major and the column-major workers)
template <typename T, bool UPDATE, bool PROF>
__device__ __forceinline__ void df_coef_warp(T *v, const DfParams &P, double *coef, volatile int *prog,
                                             volatile int *cprog, volatile int *pprog, int wkr, int lane,
                                             int reuse_lag = kDfD + 1) {
  const int n = P.n, T_ = P.T, Wk = P.Wk;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // ---------------- coefficient warp (software-pipelined: the poll of block column s+1 is in
  // flight while the coefficients of block column s are formed).  A ring slot is reused once
  // every row warp has finished the step that last read it (prog[]); finished warps have left.
  double t_run = P.ws.scal[0], S_run = P.ws.scal[1];  // (1, 1) unless a blocked driver carries them over
#ifndef CHUB_DF_POLL2
#define CHUB_DF_POLL2 0  // two polls in flight in the coefficient warp: measured 4 % SLOWER (147 SMs hammer one L2 line)
#endif
  unsigned long long b0 = ld_relaxed_u64(P.ws.p + lane);
  unsigned long long b1 = (CHUB_DF_POLL2 && T_ > 1) ? ld_relaxed_u64(P.ws.p + kPanel + lane) : kSentinelBits;
  double ndg = P.ws.dg[lane];
  for (int pan = 0; pan < T_; ++pan) {
    const int slot = pan % kDfCoefRing;
    const int k = pan * kPanel + lane;
    long long tq = DF_PROF_T();
    // two polls in flight (this block column and the next): the chain may run at two block columns per
    // round trip of a poll
    double pk;
    if (CHUB_DF_POLL2) {
      pk = df_wait_value2(P.ws.p + k, pan + 1 < T_ ? P.ws.p + k + kPanel : nullptr, b0, b1, P.status);
      b0 = b1;
      b1 = pan + 2 < T_ ? ld_relaxed_u64(P.ws.p + k + 2 * kPanel) : kSentinelBits;
    } else {  // one poll in flight: the next block column's, issued while this one is turned into coefficients
#ifdef CHUB_EXP_WORKER_FREE  // timing experiment (results are garbage): the workers never wait for the chain
      pk = b0 != kSentinelBits ? __longlong_as_double((long long)b0) : 0.0;
#else
      pk = b0 != kSentinelBits ? __longlong_as_double((long long)b0) : df_wait_value(P.ws.p + k, P.status);
#endif
      b0 = pan + 1 < T_ ? ld_relaxed_u64(P.ws.p + k + kPanel) : kSentinelBits;
    }
    const double dg = ndg;
    if (pan + 1 < T_) ndg = P.ws.dg[k + kPanel];
    if (lane == 0) DF_PROF_ADD(6, tq);
    if (lane == 0 && wkr == 0) DF_TRACE(3, pan);
    if (pan >= kDfCoefRing) {
      // slot last used by block column pan-ring.  prog[] counts finished steps (first-generation workers:
      // that block column is read for the last time at step pan-ring+d, reuse_lag = d+1) or the next block
      // column a warp still needs (rm5 workers, reuse_lag = 1)
      const int need = pan - kDfCoefRing + reuse_lag;
      const long long t0 = clock64();
      while (__reduce_min_sync(0xffffffffu, ld_acquire_cta_smem(prog + lane)) < need) {
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
      }
    }
    double *cf = coef + (size_t)slot * 4 * kPanel;
    cf[lane] = pk;
    // p alone is enough for the residual w (and for the hand-over to the chain, which is on the
    // chain <-> worker latency loop): publish it before the scan / rsqrt of the coefficients
    __syncwarp();
    if (lane == 0) st_release_cta_smem(pprog, pan + 1);
    if (UPDATE) {
      double incl = pk * pk;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 0.0;
      const double tk = t_run + excl, tk1 = t_run + incl;
      const unsigned negmask = __ballot_sync(0xffffffffu, dg < 0.0);
      const double sgn = dg < 0.0 ? -1.0 : 1.0;
      const double rinv = rsqrt(tk * tk1);
      const double ck = sgn * tk * rinv, sigk = sgn * pk * rinv;
      cf[kPanel + lane] = ck;
      cf[2 * kPanel + lane] = sigk;
      cf[3 * kPanel + lane] = fabs(dg) * tk1 * rinv;
      __syncwarp();
      if (lane == 0) st_release_cta_smem(cprog, pan + 1);
      if (wkr == pan % Wk) {  // one worker per panel: coefficients for the rectangles below a diagonal block
        P.ws.c[k] = ck;
        P.ws.sg[k] = sigk;
      }
      if (wkr == pan % Wk && k < n) {  // ... and rotg's z_k in v[k]
        const double Sk = (__popc(negmask & ((1u << lane) - 1u)) & 1) ? -S_run : S_run;
        const double nuk = Sk * dg * pk / sqrt(tk);
        const double sk = Sk * sgn * pk / sqrt(tk1);
        v[k] = (T)rotg_z<double>(dg, nuk, ck, sk);
      }
      t_run = __shfl_sync(0xffffffffu, tk1, 31);
      if (__popc(negmask) & 1) S_run = -S_run;
    } else {
      __syncwarp();
      if (lane == 0) st_release_cta_smem(cprog, pan + 1);
    }
    if (lane == 0 && wkr == 0) DF_TRACE(4, pan);
    if (lane == 0) DF_PROF_ADD(7, tq);
  }
  if (UPDATE && wkr == 0 && lane == 0) { P.ws.scal[4] = t_run; P.ws.scal[5] = S_run; }
  if (P.prof && lane == 0) {
    P.prof[(size_t)blockIdx.x * 16 + 6] = pacc[6];
    P.prof[(size_t)blockIdx.x * 16 + 7] = pacc[7];
  }
}

// ---- worker CTAs
template <typename T, bool UPDATE, bool PROF>
__device__ void df_worker_rm(T *L, T *v, const DfParams &P, const CUtensorMap *tmap8,
                             unsigned char *smem) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC;
  constexpr int CPT = RmGeom<T>::CPT, WTILE = RmGeom<T>::WTILE;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int wkr = blockIdx.x - 1;
  const int n = P.n, T_ = P.T, Wk = P.Wk, nrw = P.nw;  // nrw = row warps (2 per 16-row block)
  const int64_t ld = P.ld;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [nrw][kDfStages]
  unsigned long long *cfull = full + kDfStages * 2 * kRmMaxQ;               // [kDfCoefRing]
  volatile int *prog = reinterpret_cast<volatile int *>(cfull + kDfCoefRing);  // [32] steps finished per row warp
  volatile int *cprog = prog + 32;  // block columns whose coefficients (c, sigma, dn) are in the ring
  volatile int *pprog = prog + 33;  // block columns whose p is in the ring (always >= *cprog)
  double *coef = reinterpret_cast<double *>(smem + 2048);                   // [kDfCoefRing][4][32]
  unsigned char *stages = smem + 2048 + kDfCoefRing * 4 * kPanel * 8;       // 1024-aligned

  if (tid == 0) {
    for (int i = 0; i < kDfStages * nrw; ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < kDfCoefRing; ++i) mbar_init(&cfull[i], 32);
    fence_mbar_init();
  }
  if (tid < 32) prog[tid] = tid < nrw ? 0 : 0x7fffffff;
  if (tid == 0) { *cprog = 0; *pprog = 0; }
  __syncthreads();
  if (warp > nrw) return;

  if (warp == nrw) {
    df_coef_warp<T, UPDATE, PROF>(v, P, coef, prog, cprog, pprog, wkr, lane);
    return;
  }

  // ---------------- row warps: warp = 8 consecutive rows, lane = (segment g, row r8)
  const int g = lane >> 3, r8 = lane & 7;
  const int64_t row0 = (((int64_t)wkr + (int64_t)(warp >> 1) * Wk) << 4) + ((warp & 1) << 3);
  if (row0 >= n) {  // no rows here: leave (prog = "done" so that nobody waits for this warp)
    if (lane == 0) prog[warp] = 0x7fffffff;
    return;
  }
  const int b = (int)(row0 >> 5);
  const int last_step = min(T_ + kDfD - 1, b + kDfD);  // my rows are complete after this step
  const int64_t row = row0 + r8;
  double w = row < n ? P.ws.v0[row] : 0.0;
  const int64_t row0_next = (((int64_t)wkr + (int64_t)((warp + 1) >> 1) * Wk) << 4) + (((warp + 1) & 1) << 3);
  const bool prof_me = PROF && P.prof && lane == 0 && (warp + 1 >= nrw || row0_next >= n);
  // my 8 columns live in half `hh`, chunks cb .. cb+CPT-1 of the 128-byte box row
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  unsigned long long *myfull = full + warp * kDfStages;
  unsigned char *mytile = stages + (size_t)warp * (kDfStages * WTILE);  // [kDfStages][WTILE]
  const unsigned my_off = hh * 1024 + r8 * 128;

  auto issue_load = [&](int s, int st) {  // lane 0 only; st = s % kDfStages
    if (s > last_step) return;
    void *bar = &myfull[st];
    const int pan = df_panel_at<kDfD>(b, s);
    if (pan < 0) { mbar_arrive(bar); return; }
    unsigned char *tile = mytile + st * WTILE;
    const int c0 = pan * kPanel;
    if (HALVES == 2 && c0 + BOXC < n) {
      mbar_arrive_expect_tx(bar, 2048u);
      tma_load_2d(tile, tmap8, c0, (int)row0, bar);
      tma_load_2d(tile + 1024, tmap8, c0 + BOXC, (int)row0, bar);
    } else {
      mbar_arrive_expect_tx(bar, 1024u);
      tma_load_2d(tile, tmap8, c0, (int)row0, bar);
    }
  };

  // One block column of my 8 rows: tile in shared memory -> updated tile (or, for the ragged
  // diagonal tile, straight to global memory).  Returns true if the tile has to be TMA-stored.
  auto do_step = [&](int s, int st) -> bool {
    const int pan = df_panel_at<kDfD>(b, s);
    if (pan < 0) return false;
    unsigned char *rowp = mytile + st * WTILE + my_off;
    const double *cf = coef + (size_t)(pan % kDfCoefRing) * 4 * kPanel + g * 8;
    const int64_t c0 = (int64_t)pan * kPanel + g * 8;  // my first column
    const bool diag_step = (pan == b);
    int4 raw[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i)
      raw[i] = *reinterpret_cast<const int4 *>(rowp + (((cb + i) ^ r8) << 4));
    T *x = reinterpret_cast<T *>(raw);
    double pc[8], cc[8], sc[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 a = *reinterpret_cast<const double2 *>(cf + j);
      pc[j] = a.x; pc[j + 1] = a.y;
    }
    double lo[8], pre[8];  // pre[j] = sum_{m<j} lo_m p_m within my segment
    double run = 0.0;
    if (!diag_step) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        lo[j] = (double)x[j];
        pre[j] = run;
        run = fma(lo[j], pc[j], run);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        lo[j] = (double)x[j];
        pre[j] = run;
        if (c0 + j < row) run = fma(lo[j], pc[j], run);
      }
    }
    // exclusive scan of the four segment totals (lanes g*8 + r8, same r8)
    double incl = run;
    double y = __shfl_up_sync(0xffffffffu, incl, 8);
    if (g >= 1) incl += y;
    y = __shfl_up_sync(0xffffffffu, incl, 16);
    if (g >= 2) incl += y;
    const double total = __shfl_sync(0xffffffffu, incl, 24 + r8);
    const double wseg = w - (incl - run);
    w -= total;
    // hand the residual over to the chain after the last far-left block column -- as early as
    // possible: this store is on the chain <-> worker latency loop
    if (s == b - kDfD - 1 && g == 0 && row < n) {
      st_relaxed_f64(P.ws.w + row, df_sanitize(w));
      if (lane == 0 && (row0 & 31) == 0) DF_TRACE(0, b);
    }
    if (UPDATE) {
      // c, sigma (and dn) of this block column: normally long there; only the hand-over step runs ahead
      if (ld_acquire_cta_smem(cprog) <= pan) {
        const long long t0 = clock64();
        while (ld_acquire_cta_smem(cprog) <= pan) {
          if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 c = *reinterpret_cast<const double2 *>(cf + kPanel + j);
        const double2 d = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + j);
        cc[j] = c.x; cc[j + 1] = c.y; sc[j] = d.x; sc[j + 1] = d.y;
      }
      if (!diag_step) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (T)fma(cc[j], lo[j], sc[j] * (wseg - pre[j]));
#pragma unroll
        for (int i = 0; i < CPT; ++i)
          *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
        return true;
      }
      if (row < n) {  // the diagonal tile is ragged: written straight to global memory
        T *gp = L + row * ld + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t col = c0 + j;
          if (col < row) gp[j] = (T)fma(cc[j], lo[j], sc[j] * (wseg - pre[j]));
          else if (col == row) gp[j] = (T)cf[3 * kPanel + j];
        }
      }
    }
    return false;
  };

  auto issue_store = [&](int s, int st) {  // lane 0 only
    const int c0 = df_panel_at<kDfD>(b, s) * kPanel;
    unsigned char *tile = mytile + st * WTILE;
    tma_store_2d(tmap8, c0, (int)row0, tile);
    if (HALVES == 2 && c0 + BOXC < n) tma_store_2d(tmap8, c0 + BOXC, (int)row0, tile + 1024);
  };

  // Two block columns (steps 2S, 2S+1) per iteration: the fixed per-step costs (waits, TMA issue,
  // proxy fence, commit) are paid once per 64 columns.  Stages: pair S uses (2S) % 6, (2S+1) % 6;
  // loads run two pairs ahead (issued after the compute phase, when the store of the pair whose
  // stages they reuse has drained).
  static_assert(kDfStages == 6, "the pair schedule assumes a ring of 6 stages");
#ifndef CHUB_DF_AHEAD
#define CHUB_DF_AHEAD 2  // pairs of block columns a row warp keeps in flight beyond the one it computes
#endif
  if (elect_one()) {
    issue_load(0, 0);
    issue_load(1, 1);
    if (CHUB_DF_AHEAD >= 2) {
      issue_load(2, 2);
      issue_load(3, 3);
    }
  }
  __syncwarp();
  const long long t_wk0 = DF_PROF_T();
  int st = 0;        // (2S) % 6
  unsigned ph = 0;   // ((2S) / 6) & 1
  for (int s = 0; s <= last_step; s += 2) {
    long long tp = DF_PROF_T();
    const int st1 = st + 1;                             // stage of step s+1 (st is even: no wrap)
    const int stn = st + 2 == kDfStages ? 0 : st + 2;   // stages of the next pair
    const int stnn = st >= 2 ? st - 2 : st + 4;         // stages of pair S+2 (= those of pair S-1)
    const bool two = s + 1 <= last_step;
    mbar_wait(&myfull[st], ph, P.status);
    if (two) mbar_wait(&myfull[st1], ph, P.status);
    if (prof_me) DF_PROF_ADD(1, tp);
    tp = DF_PROF_T();
    // Step s is gated on p_s (in the ring): p_s published implies that the chain has consumed
    // everything step s overwrites.  Each step waits for its own p only, so that the hand-over step
    // (on the chain <-> worker latency loop) never waits for the block column after it.
    auto wait_p = [&](int step) {
      const int need = min(step, T_ - 1) + 1;
      if (ld_acquire_cta_smem(pprog) < need) {
        const long long t0 = clock64();
        while (ld_acquire_cta_smem(pprog) < need) {
          if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
        }
      }
    };
    wait_p(s);
    if (prof_me) DF_PROF_ADD(4, tp);
    if (prof_me && wkr == 0 && s < T_) DF_TRACE(5, s);
    tp = DF_PROF_T();
    const bool store0 = do_step(s, st);
    if (two) wait_p(s + 1);
    const bool store1 = two ? do_step(s + 1, st1) : false;
    if (prof_me) DF_PROF_ADD(3, tp);
    if (prof_me && wkr == 0 && s < T_) DF_TRACE(6, s);
    tp = DF_PROF_T();
    // The ring holds the pair being processed and two pairs of loads in flight.  Pair S+2 reuses the
    // stages of pair S-1, whose TMA store has had this whole compute phase to drain.
    if (elect_one()) {
      if (CHUB_DF_AHEAD >= 2) {
        if (s >= 2) bulk_wait_read<0>();
        issue_load(s + 4, stnn);
        issue_load(s + 5, stnn + 1);
      } else {  // one pair ahead: the next pair's stages held pair S-2, whose store is two commits back
        if (s >= 2) bulk_wait_read<1>();
        issue_load(s + 2, stn);
        issue_load(s + 3, stn + 1);
      }
    }
    if (prof_me) DF_PROF_ADD(2, tp);
    tp = DF_PROF_T();

    fence_proxy_async();  // my generic-proxy writes -> visible to the TMA (async proxy) store
    __syncwarp();
    const bool st0 = __any_sync(0xffffffffu, store0), st1b = __any_sync(0xffffffffu, store1);
    if (elect_one()) {
      // progress counter for the coefficient warp (ring-slot reuse).  Every lane's coefficient reads of
      // this pair are complete (their values went into the tile writes the fence above ordered, and the
      // __syncwarp orders the lanes), so a plain store is enough: no second MEMBAR on the row warps' path.
      prog[warp] = s + 1 >= last_step ? 0x7fffffff : s + 2;
      if (st0) issue_store(s, st);
      if (st1b) issue_store(s + 1, st1);
      bulk_commit();
    }
    __syncwarp();
    st = stn;
    if (st == 0) ph ^= 1u;
    if (prof_me) DF_PROF_ADD(5, tp);
  }
  if (prof_me) DF_PROF_ADD(0, t_wk0);
  if (elect_one()) bulk_wait_all<0>();
  if (prof_me) {
    for (int i = 0; i < 6; ++i) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_worker_rm_smem(int nrw) {
  return 2048 + (size_t)kDfCoefRing * 4 * kPanel * 8 + (size_t)kDfStages * nrw * RmGeom<T>::WTILE;
}

} // namespace chub
#include "dataflow_rm5.cuh"
#ifdef CHUB_CHAIN_V4
#include "dataflow_chain_v4.cuh"
#endif
#ifndef CHUB_CHAIN_V6
#define CHUB_CHAIN_V6 1  // 1: barrier-free chain CTA (dataflow_chain_v6.cuh); 0: the barrier chain df_chain_rm (A/B)
#endif
#include "dataflow_chain_v6.cuh"
#include "dataflow_chain_v7.cuh"
namespace chub {

// V5: second-generation row warps (dataflow_rm5.cuh); false = the first generation above (CHUB_RM5=0, A/B only)
template <typename T, bool UPDATE, bool PROF, bool V5>
__global__ void __launch_bounds__(kRmMaxThreads, 1)
df_kernel_rm(T *L, T *v, DfParams P, const __grid_constant__ CUtensorMap tmap8,
             const __grid_constant__ CUtensorMap tmap32, const __grid_constant__ CUtensorMap tmap3) {
  extern __shared__ __align__(1024) unsigned char df_smem[];
  if (*P.status != 0) return;  // check_diag failed: L and v stay untouched
  if (blockIdx.x == 0) {
#ifndef CHUB_EXP_NO_CHAIN
#ifdef CHUB_CHAIN_V4
    if (threadIdx.x < 256) df_chain_v4<T, CHUB_ROW_MAJOR, PROF>(P, &tmap32, df_smem);
#elif CHUB_CHAIN_V7
    if (threadIdx.x < kC7Threads) df_chain_v7<T, CHUB_ROW_MAJOR, PROF>(P, &tmap32, df_smem);
#elif CHUB_CHAIN_V6
    if (threadIdx.x < kC6ThreadsSplit) df_chain_v6<T, CHUB_ROW_MAJOR, PROF, false>(P, &tmap32, df_smem);
#else
    if (threadIdx.x < 256) df_chain_rm<T, CHUB_ROW_MAJOR, PROF>(P, &tmap32, df_smem);
#endif
#endif
  } else {
#ifndef CHUB_EXP_NO_WORKERS
    if (V5) df_worker_rm5<T, UPDATE, PROF>(L, v, P, &tmap8, &tmap3, df_smem);
    else df_worker_rm<T, UPDATE, PROF>(L, v, P, &tmap8, df_smem);
#endif
  }
}

// =====================================================================================
// Rectangle below a diagonal block (row-major): factors beyond the persistent kernel's row capacity
// =====================================================================================
// The persistent kernel keeps every row in flight at once, so one launch covers at most 147 x 14 x 8 =
// 16464 rows.  Larger factors are cut into diagonal blocks of at most that order (blocked driver,
// api.cu): the persistent kernel updates diagonal block J (its columns' (p, c, sigma) end up in global
// memory), and this kernel applies those columns to ALL rows below the block -- the part of the
// right-looking update that has no sequential dependency at all:
//     L_ik <- c_k L_ik + sigma_k w_i ,   w_i <- w_i - L_ik p_k      (k over the block's columns)
// leaving the residual w_i in ws.w for the next diagonal block.  A warp takes 8 rows and walks the block's
// columns in 32-column tiles through a ring of TMA boxes, the block column's coefficients riding along;
// same lane mapping and arithmetic as the persistent kernel's workers.  UPDATE = false: residual only
// (forward substitution of the downdate).
constexpr int kRectStages = 4;
constexpr int kRectWarps = 14;

template <typename T>
struct RectGeom {
  static constexpr int STAGE = RmGeom<T>::WTILE + 1024;  // tile + (p, c, sigma)[32] doubles (768 B, padded)
};

template <typename T, bool UPDATE>
__global__ void __launch_bounds__(kRectWarps * 32, 1)
rect_apply_rm_kernel(int n, int row_begin, int col_begin, int ncols, LargeWs ws, const int32_t *status,
                     const __grid_constant__ CUtensorMap tmap8) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, CPT = RmGeom<T>::CPT;
  constexpr int WTILE = RmGeom<T>::WTILE, STAGE = RectGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char rc_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(rc_smem) + warp * kRectStages;
  unsigned char *mystage = rc_smem + 1024 + (size_t)warp * kRectStages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRectWarps * kRectStages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(rc_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int g = lane >> 3, r8 = lane & 7;
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  const unsigned my_off = hh * 1024 + r8 * 128;
  const int nsteps = ncols / kPanel;
  const int ngroups = (n - row_begin + 7) / 8;
  unsigned it = 0;  // tiles this warp has consumed so far (ring position / mbarrier phase)
  for (int grp = (int)blockIdx.x * kRectWarps + warp; grp < ngroups; grp += (int)gridDim.x * kRectWarps) {
    const int row0 = row_begin + grp * 8;
    const int row = row0 + r8;
    double w = row < n ? ws.w[row] : 0.0;

    auto issue_load = [&](int s) {  // elected lane; stage / phase follow the warp's running tile count
      if (s >= nsteps) return;
      const unsigned st = (it + (unsigned)s) % kRectStages;
      void *bar = &full[st];
      unsigned char *stg = mystage + (size_t)st * STAGE;
      const int c0 = col_begin + s * kPanel;
      mbar_arrive_expect_tx(bar, (unsigned)WTILE + (UPDATE ? 768u : 256u));
      tma_load_2d(stg, &tmap8, c0, row0, bar);
      if (HALVES == 2) tma_load_2d(stg + 1024, &tmap8, c0 + BOXC, row0, bar);
      bulk_g2s(stg + WTILE, ws.p + c0, 256u, bar);
      if (UPDATE) {
        bulk_g2s(stg + WTILE + 256, ws.c + c0, 256u, bar);
        bulk_g2s(stg + WTILE + 512, ws.sg + c0, 256u, bar);
      }
    };

    if (elect_one()) {
      bulk_wait_read<0>();  // (the previous group's last stores have left their stages)
      for (int s = 0; s < kRectStages - 1; ++s) issue_load(s);
    }
    __syncwarp();
    for (int s = 0; s < nsteps; ++s) {
      const unsigned st = (it + (unsigned)s) % kRectStages;
      const unsigned ph = ((it + (unsigned)s) / kRectStages) & 1u;
      mbar_wait(&full[st], ph, const_cast<int32_t *>(status));
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
      if (UPDATE) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 c = *reinterpret_cast<const double2 *>(cf + kPanel + j);
          const double2 d = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + j);
          x[j] = (T)fma(c.x, lo[j], d.x * (wseg - pre[j]));
          x[j + 1] = (T)fma(c.y, lo[j + 1], d.y * (wseg - pre[j + 1]));
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
        fence_proxy_async();
      }
      __syncwarp();
      if (elect_one()) {
        if (UPDATE) {
          const int c0 = col_begin + s * kPanel;
          tma_store_2d(&tmap8, c0, row0, stg);
          if (HALVES == 2) tma_store_2d(&tmap8, c0 + BOXC, row0, stg + 1024);
          bulk_commit();
          bulk_wait_read<1>();  // the store of step s-1 has left the stage that step s+3 loads into
        }
        issue_load(s + kRectStages - 1);
      }
      __syncwarp();
    }
    it += (unsigned)nsteps;
    if (g == 0 && row < n) ws.w[row] = w;
  }
  if (elect_one()) bulk_wait_all<0>();
}

template <typename T>
inline size_t rect_apply_rm_smem() { return 1024 + (size_t)kRectWarps * kRectStages * RectGeom<T>::STAGE; }

// The blocked driver's first kernel: w <- v for every row, check_diag over the WHOLE diagonal (the
// persistent kernel's pre-pass only sees its own block), (t, S) <- (1, 1).
template <typename T, int LAYOUT>
__global__ void blocked_init_kernel(const T *L, int n, int64_t ld, const T *v, LargeWs ws, int flags, int32_t *status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pad = (n + 31) & ~31;
  if (i < n_pad) ws.w[i] = i < n ? df_sanitize((double)v[i]) : 0.0;
  if (i < n && (flags & CHUB_FLAG_CHECK_DIAG) && L[idx<LAYOUT>(i, i, ld)] == T(0)) atomicMax(status, CHUB_STATUS_ZERO_DIAG);
  if (i == 0) { ws.scal[4] = 1.0; ws.scal[5] = 1.0; }
}

// =====================================================================================
// Column-major (the reference's recommended layout, _seeger.py:65-68): native tensor-map workers
// =====================================================================================
// Rows are the contiguous dimension, so a warp owns a whole block row (32 rows, lane = row) and moves one
// [32 rows x 32 columns] box per step with ONE TMA instruction (256-byte runs, no swizzle: lane = row reads
// 8 bytes per column, conflict free).  The row's residual chain over the 32 columns is cut into four
// independent 8-column segments (prefix sums in registers; this kernel has few threads per CTA, so each
// thread may keep the whole row segment of the tile), combined by three subtractions -- no shuffles.  Per
// byte this needs a quarter of the instructions of the row-major mapping.  Same dataflow otherwise: same
// chain CTA (band tiles read through a column-major box), same coefficient warp, same delay of d steps
// inside the band, ragged diagonal tile written with predicated global stores (coalesced: for a fixed
// column the 32 lanes are 32 consecutive rows).
#ifndef CHUB_CM_STAGES
#define CHUB_CM_STAGES 4
#endif
constexpr int kCmStages = CHUB_CM_STAGES;  // tile ring per row warp; loads run kCmStages - 2 steps ahead
constexpr int kCmMaxQ = 6;  // row warps (block rows) per worker: n <= 147 * 6 * 32 = 28224 (shared memory permitting)

template <typename T>
struct CmGeom {
  static constexpr int TILE = kPanel * kPanel * (int)sizeof(T);  // 8 KB / 4 KB
};

template <typename T, bool UPDATE, bool PROF>
__device__ void df_worker_cm(T *L, T *v, const DfParams &P, const CUtensorMap *tmapc, unsigned char *smem) {
  constexpr int TILE = CmGeom<T>::TILE;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wkr = blockIdx.x - 1;
  const int n = P.n, T_ = P.T, Wk = P.Wk, nrw = P.nw;
  const int64_t ld = P.ld;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);      // [nrw][kCmStages]
  volatile int *prog = reinterpret_cast<volatile int *>(smem + 256);             // [32] steps finished per row warp
  volatile int *cprog = prog + 32;
  volatile int *pprog = prog + 33;
  double *coef = reinterpret_cast<double *>(smem + 2048);                        // [kDfCoefRing][4][32]
  unsigned char *stages = smem + 2048 + kDfCoefRing * 4 * kPanel * 8;            // [nrw][kCmStages][TILE]

  if (tid == 0) {
    for (int i = 0; i < kCmStages * nrw; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  if (tid < 32) prog[tid] = tid < nrw ? 0 : 0x7fffffff;
  if (tid == 0) { *cprog = 0; *pprog = 0; }
  __syncthreads();
  if (warp > nrw + 1) return;
  if (warp == nrw + 1) {  // relay warp: p_s from global memory into the ring (split from the coefficient warp)
    df_relay_warp<PROF>(P, coef, prog, pprog, wkr, lane, kDfD + 1);
    return;
  }
  if (warp == nrw) {
    df_coef_warp5<T, UPDATE, PROF>(v, P, coef, cprog, pprog, wkr, lane);
    return;
  }

  // ---------------- row warps: warp = block row b (32 rows), lane = row
  const int b = wkr + warp * Wk;
  const int64_t row0 = (int64_t)b * kPanel;
  if (row0 >= n) {
    if (lane == 0) prog[warp] = 0x7fffffff;
    return;
  }
  const int64_t row = row0 + lane;
  const bool valid = row < n;
  const int last_step = min(T_ + kDfD - 1, b + kDfD);
  double w = valid ? P.ws.v0[row] : 0.0;
  const bool prof_me = PROF && P.prof && lane == 0 && (warp + 1 >= nrw || (int64_t)(b + Wk) * kPanel >= n);
  unsigned long long *myfull = full + warp * kCmStages;
  unsigned char *mytile = stages + (size_t)warp * (kCmStages * TILE);

  auto issue_load = [&](int s) {  // elected lane
    if (s > last_step) return;
    void *bar = &myfull[s % kCmStages];
    const int pan = df_panel_at<kDfD>(b, s);
    if (pan < 0) { mbar_arrive(bar); return; }
    mbar_arrive_expect_tx(bar, (unsigned)TILE);
    tma_load_2d(mytile + (s % kCmStages) * TILE, tmapc, (int)row0, pan * kPanel, bar);
  };

  auto wait_flag = [&](volatile int *flag, int need) {
    if (ld_acquire_cta_smem(flag) < need) {
      const long long t0 = clock64();
      while (ld_acquire_cta_smem(flag) < need) {
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
      }
    }
  };

  if (elect_one()) {
    for (int s = 0; s < kCmStages - 2; ++s) issue_load(s);
  }
  __syncwarp();
  const long long t_wk0 = DF_PROF_T();
  for (int s = 0; s <= last_step; ++s) {
    long long tp = DF_PROF_T();
    if (elect_one()) {
      bulk_wait_read<1>();  // the store of step s-2 has left its stage (= the stage of step s + kCmStages - 2)
      issue_load(s + kCmStages - 2);
    }
    __syncwarp();
    if (prof_me) DF_PROF_ADD(2, tp);
    tp = DF_PROF_T();
    mbar_wait(&myfull[s % kCmStages], (unsigned)((s / kCmStages) & 1), P.status);
    if (prof_me) DF_PROF_ADD(1, tp);
    tp = DF_PROF_T();
    // step s is gated on p_s: p_s published implies that the chain has consumed everything step s overwrites
    wait_flag(pprog, min(s, T_ - 1) + 1);
    if (prof_me) DF_PROF_ADD(4, tp);
    tp = DF_PROF_T();
    const int pan = df_panel_at<kDfD>(b, s);
    bool do_store = false;
    if (pan >= 0) {
      T *tile = reinterpret_cast<T *>(mytile + (s % kCmStages) * TILE) + lane;  // element (lane, k) at tile[k * 32]
      const double *cf = coef + (size_t)(pan % kDfCoefRing) * 4 * kPanel;
      const bool diag = (pan == b);  // columns c0 + k, c0 = row0: k < lane below, k == lane on the diagonal
      double lo[kPanel], pre[kPanel], tot[4];
#pragma unroll
      for (int k = 0; k < kPanel; ++k) lo[k] = (double)tile[k * kPanel];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int k = 8 * q + j;
          const double2 pp = *reinterpret_cast<const double2 *>(cf + k);
          pre[k] = run;
          if (!diag || k < lane) run = fma(lo[k], pp.x, run);
          pre[k + 1] = run;
          if (!diag || k + 1 < lane) run = fma(lo[k + 1], pp.y, run);
        }
        tot[q] = run;
      }
      double wq[4];
      wq[0] = w;
      wq[1] = wq[0] - tot[0];
      wq[2] = wq[1] - tot[1];
      wq[3] = wq[2] - tot[2];
      w = wq[3] - tot[3];
      // hand the residual over to the chain after the last far-left block column
      if (s == b - kDfD - 1 && valid) {
        st_relaxed_f64(P.ws.w + row, df_sanitize(w));
        if (lane == 0) DF_TRACE(0, b);
      }
      if (UPDATE) {
        wait_flag(cprog, pan + 1);
        if (!diag) {
#pragma unroll
          for (int k = 0; k < kPanel; k += 2) {
            const double2 c2 = *reinterpret_cast<const double2 *>(cf + kPanel + k);
            const double2 s2 = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + k);
            tile[k * kPanel] = (T)fma(c2.x, lo[k], s2.x * (wq[k >> 3] - pre[k]));
            tile[(k + 1) * kPanel] = (T)fma(c2.y, lo[k + 1], s2.y * (wq[k >> 3] - pre[k + 1]));
          }
          do_store = true;
        } else if (valid) {  // ragged diagonal tile: straight to global memory, nothing above the diagonal
          T *gp = L + row + (int64_t)row0 * ld;  // (row, row0 + k) at gp[k * ld]
#pragma unroll
          for (int k = 0; k < kPanel; ++k) {
            if (k < lane) gp[(int64_t)k * ld] = (T)fma(cf[kPanel + k], lo[k], cf[2 * kPanel + k] * (wq[k >> 3] - pre[k]));
            else if (k == lane) gp[(int64_t)k * ld] = (T)cf[3 * kPanel + k];
          }
        }
      }
    }
    if (prof_me) DF_PROF_ADD(3, tp);
    tp = DF_PROF_T();
    fence_proxy_async();  // my generic-proxy writes -> visible to the TMA (async proxy) store
    __syncwarp();
    if (elect_one()) {
      prog[warp] = s >= last_step ? 0x7fffffff : s + 1;
      if (do_store) tma_store_2d(tmapc, (int)row0, pan * kPanel, mytile + (s % kCmStages) * TILE);
      bulk_commit();
    }
    __syncwarp();
    if (prof_me) DF_PROF_ADD(5, tp);
  }
  if (prof_me) DF_PROF_ADD(0, t_wk0);
  if (elect_one()) bulk_wait_all<0>();
  if (prof_me) {
    for (int i = 0; i < 6; ++i) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_worker_cm_smem(int nrw) {
  return 2048 + (size_t)kDfCoefRing * 4 * kPanel * 8 + (size_t)kCmStages * nrw * CmGeom<T>::TILE;
}

// ---- column-major workers, second generation: TWO warps per block row.
// What the one-warp-per-block-row worker is bound by (gpurun r03m, n = 16384): 1400 of its 2050 cycles per step
// are the tile arithmetic of ONE warp (32 shared-memory loads, four 8-long dependent FMA chains, 32 updates and
// stores per lane, nothing else resident on its scheduler to hide the latencies) -- and that warp is on the
// chain <-> worker hand-over loop.  Here warps 2q and 2q+1 share block row q's tile: warp h takes columns
// [16 h, 16 h + 16).  The prefix products do not depend on the residual, so both halves run concurrently; the
// four 8-column segment totals cross through shared memory (one named barrier of 64 threads), every warp then
// forms w_i - (prefix) exactly as the one-warp kernel does (same association: bit-identical results).  Warp 2q's
// elected lane issues all TMA loads and stores of the pair (bulk async-groups are per thread).
// H = warps per block row: 2 (any Q <= kCmMaxQ) or 4 (Q <= kCm4MaxQ: the thread count caps the registers).
constexpr int kCm4MaxQ = 4;

template <int H> struct Cm2Stages { static constexpr int N = H == 4 ? 6 : kCmStages; };  // (H = 4: Q <= 4, room for six)

template <typename T, bool UPDATE, bool PROF, int kCmH>
__device__ void df_worker_cm2(T *L, T *v, const DfParams &P, const CUtensorMap *tmapc, unsigned char *smem) {
  constexpr int TILE = CmGeom<T>::TILE;
  constexpr int kCmCW = kPanel / kCmH;  // columns per warp and tile
  constexpr int kCmSeg = kCmCW / 8;     // 8-column segments per warp
  constexpr int NS = Cm2Stages<kCmH>::N;  // tile ring per block row; loads run NS - 2 steps ahead
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wkr = blockIdx.x - 1;
  const int n = P.n, T_ = P.T, Wk = P.Wk, nrw = P.nw;  // nrw = block rows (warp pairs) per worker
  const int64_t ld = P.ld;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);      // [nrw][NS]
  volatile int *prog = reinterpret_cast<volatile int *>(smem + 256);             // [32] steps finished per block row
  volatile int *cprog = prog + 32;
  volatile int *pprog = prog + 33;
  double *coef = reinterpret_cast<double *>(smem + 2048);                        // [kDfCoefRing][4][32]
  unsigned char *stages = smem + 2048 + kDfCoefRing * 4 * kPanel * 8;            // [nrw][NS][TILE]
  double *xch = reinterpret_cast<double *>(stages + (size_t)nrw * NS * TILE);  // [nrw][2][4][32] segment totals

  if (tid == 0) {
    for (int i = 0; i < NS * nrw; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  if (tid < 32) prog[tid] = tid < nrw ? 0 : 0x7fffffff;
  if (tid == 0) { *cprog = 0; *pprog = 0; }
  __syncthreads();
  if (warp > kCmH * nrw + 1) return;
  if (warp == kCmH * nrw + 1) {
    df_relay_warp<PROF>(P, coef, prog, pprog, wkr, lane, kDfD + 1);
    return;
  }
  if (warp == kCmH * nrw) {
    df_coef_warp5<T, UPDATE, PROF>(v, P, coef, cprog, pprog, wkr, lane);
    return;
  }

  // ---------------- row warps: pair q = block row (32 rows, lane = row), h = which 16 columns of every tile
  const int q = warp / kCmH, h = warp % kCmH;
  const int b = wkr + q * Wk;
  const int64_t row0 = (int64_t)b * kPanel;
  if (row0 >= n) {  // (both warps of the pair leave: nobody is left at the pair's barrier)
    if (lane == 0 && h == 0) prog[q] = 0x7fffffff;
    return;
  }
  const int64_t row = row0 + lane;
  const bool valid = row < n;
  const int last_step = min(T_ + kDfD - 1, b + kDfD);
  double w = valid ? P.ws.v0[row] : 0.0;
  const bool prof_me = PROF && P.prof && lane == 0 && h == 0 && (q + 1 >= nrw || (int64_t)(b + Wk) * kPanel >= n);
  unsigned long long *myfull = full + q * NS;
  unsigned char *mytile = stages + (size_t)q * (NS * TILE);
  double *myx = xch + (size_t)q * (2 * 4 * kPanel);
  const int bar_id = 1 + q;
  const int k0 = h * kCmCW;  // my first column within a tile
  unsigned nx = 0;           // tiles processed so far (parity = exchange buffer)

  auto issue_load = [&](int s) {  // elected lane of warp h = 0
    if (s > last_step) return;
    void *bar = &myfull[s % NS];
    const int pan = df_panel_at<kDfD>(b, s);
    if (pan < 0) { mbar_arrive(bar); return; }
    mbar_arrive_expect_tx(bar, (unsigned)TILE);
    tma_load_2d(mytile + (s % NS) * TILE, tmapc, (int)row0, pan * kPanel, bar);
  };
  auto wait_flag = [&](volatile int *flag, int need) {
    if (ld_acquire_cta_smem(flag) < need) {
      const long long t0 = clock64();
      while (ld_acquire_cta_smem(flag) < need) {
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
      }
    }
  };

  // The update of tile s is DEFERRED to step s + 1 (after that step's prefix products): (c, sigma) of a block
  // column leave the coefficient warp ~0.6 us after its p arrives (scan + rsqrt), and a row warp that waited for
  // them inside the step could not start the next step before then -- the wait was 40 % of the one-warp worker's
  // step and set the kernel's period (gpurun r04b).  One step later they are always there.  The deferred update
  // reloads its 16 entries per lane from the tile (still in its stage) and rebuilds the prefix products on the fly.
  int dpan = -1, dstage = 0;   // the deferred tile: block column, ring stage
  double dwq[kCmSeg];          // ... and the residual in front of each of my segments
  auto deferred_update = [&]() {  // both warps of the pair, together (they share the barrier)
    if (!UPDATE || dpan < 0) return;
    wait_flag(cprog, dpan + 1);
    T *tile = reinterpret_cast<T *>(mytile + dstage * TILE) + lane + k0 * kPanel;
    const double *cf = coef + (size_t)(dpan % kDfCoefRing) * 4 * kPanel + k0;
    const bool diag = (dpan == b);
    if (!diag) {
#pragma unroll
      for (int g = 0; g < kCmSeg; ++g) {
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int k = 8 * g + j;
          const double2 pp = *reinterpret_cast<const double2 *>(cf + k);
          const double2 c2 = *reinterpret_cast<const double2 *>(cf + kPanel + k);
          const double2 s2 = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + k);
          const double l0 = (double)tile[k * kPanel], l1 = (double)tile[(k + 1) * kPanel];
          tile[k * kPanel] = (T)fma(c2.x, l0, s2.x * (dwq[g] - run));
          run = fma(l0, pp.x, run);
          tile[(k + 1) * kPanel] = (T)fma(c2.y, l1, s2.y * (dwq[g] - run));
          run = fma(l1, pp.y, run);
        }
      }
      fence_proxy_async();  // my generic-proxy writes -> visible to the TMA (async proxy) store
    } else if (valid) {  // ragged diagonal tile: straight to global memory, nothing above the diagonal
      T *gp = L + row + (int64_t)(row0 + k0) * ld;  // (row, row0 + k0 + k) at gp[k * ld]
#pragma unroll
      for (int g = 0; g < kCmSeg; ++g) {
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 8 * g + j;
          const double l0 = (double)tile[k * kPanel];
          if (k0 + k < lane) {
            gp[(int64_t)k * ld] = (T)fma(cf[kPanel + k], l0, cf[2 * kPanel + k] * (dwq[g] - run));
            run = fma(l0, cf[k], run);
          } else if (k0 + k == lane) {
            gp[(int64_t)k * ld] = (T)cf[3 * kPanel + k];
          }
        }
      }
    }
    named_sync(bar_id, 32 * kCmH);  // the other warp's tile writes and coefficient reads are complete too
    if (h == 0 && elect_one()) {
      if (!diag) tma_store_2d(tmapc, (int)row0, dpan * kPanel, mytile + dstage * TILE);
      bulk_commit();
    }
    __syncwarp();
    dpan = -1;
  };

  if (h == 0 && elect_one()) {
    for (int s = 0; s < NS - 2; ++s) issue_load(s);
  }
  __syncwarp();
  const long long t_wk0 = DF_PROF_T();
  for (int s = 0; s <= last_step; ++s) {
    long long tp = DF_PROF_T();
    mbar_wait(&myfull[s % NS], (unsigned)((s / NS) & 1), P.status);
    if (prof_me) DF_PROF_ADD(1, tp);
    tp = DF_PROF_T();
    // step s is gated on p_s: p_s published implies that the chain has consumed everything step s overwrites
    wait_flag(pprog, min(s, T_ - 1) + 1);
    if (prof_me) DF_PROF_ADD(4, tp);
    tp = DF_PROF_T();
    const int pan = df_panel_at<kDfD>(b, s);  // (the same for both warps of the pair: they take the same barriers)
    int npan = -1;
    double nwq[kCmSeg];
    if (pan >= 0) {
      const T *tile = reinterpret_cast<const T *>(mytile + (s % NS) * TILE) + lane + k0 * kPanel;  // (lane, k0 + k) at tile[k * 32]
      const double *cf = coef + (size_t)(pan % kDfCoefRing) * 4 * kPanel + k0;
      const bool diag = (pan == b);  // columns c0 + k, c0 = row0: k < lane below, k == lane on the diagonal
      double lo[kCmCW];
      double *xs = myx + (nx++ & 1) * (4 * kPanel) + lane;  // segment g's total of this row at xs[g * 32] (two buffers:
      // a warp may write step s+1's totals while the other one still reads step s's)
#pragma unroll
      for (int k = 0; k < kCmCW; ++k) lo[k] = (double)tile[k * kPanel];
#pragma unroll
      for (int g = 0; g < kCmSeg; ++g) {
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int k = 8 * g + j;
          const double2 pp = *reinterpret_cast<const double2 *>(cf + k);
          if (!diag || k0 + k < lane) run = fma(lo[k], pp.x, run);
          if (!diag || k0 + k + 1 < lane) run = fma(lo[k + 1], pp.y, run);
        }
        xs[(h * kCmSeg + g) * kPanel] = run;
      }
      named_sync(bar_id, 32 * kCmH);
      double wq[4];
      wq[0] = w;
      wq[1] = wq[0] - xs[0];
      wq[2] = wq[1] - xs[kPanel];
      wq[3] = wq[2] - xs[2 * kPanel];
      w = wq[3] - xs[3 * kPanel];
      // hand the residual over to the chain after the last far-left block column
      if (s == b - kDfD - 1 && valid && h == 0) {
        st_relaxed_f64(P.ws.w + row, df_sanitize(w));
        if (lane == 0) DF_TRACE(0, b);
      }
#pragma unroll
      for (int g = 0; g < kCmSeg; ++g) {  // nwq[g] = wq[h * kCmSeg + g], as selects (no local array)
        if (kCmH == 2) nwq[g] = h == 0 ? wq[g] : wq[kCmSeg + g];
        else nwq[g] = h < 2 ? (h == 0 ? wq[0] : wq[1]) : (h == 2 ? wq[2] : wq[3]);
      }
      npan = pan;
    }
    if (prof_me) DF_PROF_ADD(3, tp);
    tp = DF_PROF_T();
    if (UPDATE) {
      if (h == 0 && elect_one()) {
        // tile s+2 goes into the stage of tile s-2, whose store was committed one step ago
        bulk_wait_read<0>();
        issue_load(s + NS - 2);
      }
      __syncwarp();
      if (prof_me) DF_PROF_ADD(2, tp);
      tp = DF_PROF_T();
      deferred_update();  // tile s-1
      if (h == 0 && lane == 0) prog[q] = s;  // steps < s: all coefficient reads done (ring-slot reuse)
      if (npan >= 0) {
        dpan = npan;
        dstage = s % NS;
#pragma unroll
        for (int g = 0; g < kCmSeg; ++g) dwq[g] = nwq[g];
      }
    } else {
      // forward substitution only: nothing is written; both warps' reads of p_s are behind the barrier above
      if (h == 0 && elect_one()) {
        prog[q] = s >= last_step ? 0x7fffffff : s + 1;
        issue_load(s + NS - 2);
      }
      __syncwarp();
    }
    if (prof_me) DF_PROF_ADD(5, tp);
  }
  deferred_update();  // the last tile (the diagonal one)
  if (UPDATE && h == 0 && lane == 0) prog[q] = 0x7fffffff;
  if (prof_me) DF_PROF_ADD(0, t_wk0);
  if (h == 0 && elect_one()) bulk_wait_all<0>();
  if (prof_me) {
    for (int i = 0; i < 6; ++i) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_worker_cm2_smem(int nrw, int H) {
  const int ns = H == 4 ? Cm2Stages<4>::N : Cm2Stages<2>::N;
  return 2048 + (size_t)kDfCoefRing * 4 * kPanel * 8 + (size_t)ns * nrw * CmGeom<T>::TILE + (size_t)nrw * 2 * 4 * kPanel * 8;
}

template <int H> struct Cm2Threads {  // row warps + coefficient + relay warp
  static constexpr int MAX = 32 * (H * (H == 2 ? kCmMaxQ : kCm4MaxQ) + 2);
};

template <typename T, bool UPDATE, bool PROF, int H>
__global__ void __launch_bounds__(Cm2Threads<H>::MAX, 1)
df_kernel_cm2(T *L, T *v, DfParams P, const __grid_constant__ CUtensorMap tmapc,
              const __grid_constant__ CUtensorMap tmapcb) {
  extern __shared__ __align__(1024) unsigned char df_smem[];
  if (*P.status != 0) return;  // check_diag failed: L and v stay untouched
  if (blockIdx.x == 0) {
    // (split poll / TMA warps, as in the row-major kernel: these workers leave the registers for 288 chain threads)
#if CHUB_CHAIN_V7
    if (threadIdx.x < kC7Threads) df_chain_v7<T, CHUB_COL_MAJOR, PROF>(P, &tmapcb, df_smem);
#else
    if (threadIdx.x < kC6ThreadsSplit) df_chain_v6<T, CHUB_COL_MAJOR, PROF, false>(P, &tmapcb, df_smem);
#endif
  } else {
    df_worker_cm2<T, UPDATE, PROF, H>(L, v, P, &tmapc, df_smem);
  }
}

constexpr int kCmMaxThreads = kC6ThreadsMerged > 256 ? kC6ThreadsMerged : 256;  // chain: v6 warp roles; workers: kCmMaxQ row warps + coefficient + relay warp

template <typename T, bool UPDATE, bool PROF>
__global__ void __launch_bounds__(kCmMaxThreads, 1)
df_kernel_cm(T *L, T *v, DfParams P, const __grid_constant__ CUtensorMap tmapc,
             const __grid_constant__ CUtensorMap tmapcb) {
  extern __shared__ __align__(1024) unsigned char df_smem[];
  if (*P.status != 0) return;  // check_diag failed: L and v stay untouched
  if (blockIdx.x == 0) {
#if CHUB_CHAIN_V6
    if (threadIdx.x < kC6ThreadsMerged) df_chain_v6<T, CHUB_COL_MAJOR, PROF, true>(P, &tmapcb, df_smem);
#else
    if (threadIdx.x < 256) df_chain_rm<T, CHUB_COL_MAJOR, PROF>(P, &tmapcb, df_smem);
#endif
  } else {
    df_worker_cm<T, UPDATE, PROF>(L, v, P, &tmapc, df_smem);
  }
}

// ---- downdate sweep, column-major: one warp per block row (lane = row), tiles from the diagonal to
// column 0 through a ring of TMA boxes with the block column's (c, s, sgn) riding along.  The reverse sweep
// (_seeger_impl_cython.pyx:285-330, row-independent form of SURVEY A.3) is a 32-long dependent FMA chain per
// tile and lane; eight warps per SM hide it.
constexpr int kSwCmStages = 3;

template <typename T>
struct SwCmGeom {
  static constexpr int STAGE = CmGeom<T>::TILE + 1024;  // tile + (c, s, sgn)[32] doubles (768 B, padded)
};

template <typename T>
__global__ void __launch_bounds__(128) dd_sweep_cm_kernel(T *L, int n, int64_t ld, LargeWs ws, const int32_t *status,
                                                          const __grid_constant__ CUtensorMap tmapc) {
  constexpr int TILE = CmGeom<T>::TILE, STAGE = SwCmGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char sw_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int T_ = (n + 31) / 32;
  const int b = T_ - 1 - ((int)blockIdx.x * 4 + warp);  // longest rows first
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sw_smem) + warp * kSwCmStages;
  unsigned char *mystage = sw_smem + 1024 + (size_t)warp * kSwCmStages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * kSwCmStages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(sw_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (b < 0) return;
  const int64_t row0 = (int64_t)b * kPanel, row = row0 + lane;
  const bool valid = row < n;
  const int nsteps = b + 1;  // tiles b, b-1, ..., 0

  auto issue_load = [&](int s) {  // elected lane
    if (s >= nsteps) return;
    const int pan = b - s;
    void *bar = &full[s % kSwCmStages];
    unsigned char *stg = mystage + (size_t)(s % kSwCmStages) * STAGE;
    mbar_arrive_expect_tx(bar, (unsigned)TILE + 768u);
    tma_load_2d(stg, &tmapc, (int)row0, pan * kPanel, bar);
    bulk_g2s(stg + TILE, ws.c + pan * kPanel, 256u, bar);
    bulk_g2s(stg + TILE + 256, ws.sg + pan * kPanel, 256u, bar);
    bulk_g2s(stg + TILE + 512, ws.dn + pan * kPanel, 256u, bar);
  };

  if (elect_one()) {
    issue_load(0);
    issue_load(1);
  }
  __syncwarp();
  double x = 0.0;  // x_i carried from the tile to the right
  for (int s = 0; s < nsteps; ++s) {
    if (elect_one()) {
      bulk_wait_read<0>();  // the store of step s-1 has left its stage (= the stage of step s+2)
      issue_load(s + 2);
    }
    __syncwarp();
    mbar_wait(&full[s % kSwCmStages], (unsigned)((s / kSwCmStages) & 1), const_cast<int32_t *>(status));
    const int pan = b - s;
    unsigned char *stg = mystage + (size_t)(s % kSwCmStages) * STAGE;
    T *tile = reinterpret_cast<T *>(stg) + lane;  // element (lane, k) at tile[k * 32]
    const double *cf = reinterpret_cast<const double *>(stg + TILE);
    const bool diag = (s == 0);
    double lnew[kPanel];
#pragma unroll
    for (int k = kPanel - 1; k >= 0; --k) {
      const double lo = (double)tile[k * kPanel];
      const double c = cf[k], sn = cf[kPanel + k];
      const double ln = cf[2 * kPanel + k] * (c * lo - sn * x);
      const double xn = fma(c, x, sn * lo);
      // above the diagonal (diagonal tile, k > lane): identity, the entry's bits never reach x or L
      const bool on = !diag || k <= lane;
      lnew[k] = ln;
      x = on ? xn : x;
    }
    if (!diag) {
#pragma unroll
      for (int k = 0; k < kPanel; ++k) tile[k * kPanel] = (T)lnew[k];
    } else if (valid) {
      T *gp = L + row + (int64_t)row0 * ld;
#pragma unroll
      for (int k = 0; k < kPanel; ++k)
        if (k <= lane) gp[(int64_t)k * ld] = (T)lnew[k];
    }
    fence_proxy_async();
    __syncwarp();
    if (elect_one()) {
      if (!diag) tma_store_2d(&tmapc, (int)row0, pan * kPanel, stg);
      bulk_commit();
    }
    __syncwarp();
  }
  if (elect_one()) bulk_wait_all<0>();
}

// ---- the same sweep with the block rows dealt so that every SM moves the same number of tiles.  Block row b
// has b + 1 tiles; dd_sweep_cm_kernel gives CTA i the four longest remaining rows, so the first CTA (2040 of the
// 131 k tiles at n = 16384) sets the kernel's time at the bandwidth ONE SM can pull: 552 us for 2.15 GB
// (profiles/r04j launch list).  Here one CTA per SM (eight warps), rows sorted by length and dealt in a snake:
// round j goes c = 0 .. G-1 for even j and back for odd j, warp w takes rounds w, w + 8, ...
constexpr int kSwCm2Warps = 8;
constexpr int kSwCm2StagesTotal = 24;  // ring stages per CTA, shared out over the warps that have rows (ns each):
// a warp streams one block row and is bound by (tiles in flight) / (memory latency), so with few rows per SM
// (n = 16384: 3.5) each of them gets a deep ring; ns >= 4: loads run ns - 2 steps ahead and land in the stage
// stored from two steps ago; ns = 3: two steps ahead, behind a wait for the previous step's store

template <typename T>
__global__ void __launch_bounds__(kSwCm2Warps * 32, 1)
dd_sweep_cm2_kernel(T *L, int n, int64_t ld, LargeWs ws, const int32_t *status, int ns,
                    const __grid_constant__ CUtensorMap tmapc) {
  constexpr int TILE = CmGeom<T>::TILE, STAGE = SwCmGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char sw_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int T_ = (n + 31) / 32;
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sw_smem) + warp * ns;
  unsigned char *mystage = sw_smem + 1024 + (size_t)warp * ns * STAGE;
  const bool deep = ns >= 4;
  const int pf = deep ? ns - 2 : ns - 1;  // prefetch distance in steps
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSwCm2StagesTotal; ++i) mbar_init(reinterpret_cast<unsigned long long *>(sw_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  unsigned it = 0;  // tiles this warp has consumed so far (ring position / mbarrier phase)
  const int R = (T_ + G - 1) / G;  // rounds
  for (int j = warp; j < R; j += kSwCm2Warps) {
    const int pos = j * G + ((j & 1) ? G - 1 - c : c);  // position in the list of rows sorted by length
    if (pos >= T_) continue;  // (last round only)
    const int b = T_ - 1 - pos;
    const int64_t row0 = (int64_t)b * kPanel, row = row0 + lane;
    const bool valid = row < n;
    const int nsteps = b + 1;  // tiles b, b-1, ..., 0

    auto issue_load = [&](int s) {  // elected lane
      if (s >= nsteps) return;
      const int pan = b - s;
      const unsigned st = (it + (unsigned)s) % (unsigned)ns;
      void *bar = &full[st];
      unsigned char *stg = mystage + (size_t)st * STAGE;
      mbar_arrive_expect_tx(bar, (unsigned)TILE + 768u);
      tma_load_2d(stg, &tmapc, (int)row0, pan * kPanel, bar);
      bulk_g2s(stg + TILE, ws.c + pan * kPanel, 256u, bar);
      bulk_g2s(stg + TILE + 256, ws.sg + pan * kPanel, 256u, bar);
      bulk_g2s(stg + TILE + 512, ws.dn + pan * kPanel, 256u, bar);
    };

    if (elect_one()) {
      bulk_wait_read<0>();  // (the previous row's last stores have left their stages)
      for (int s = 0; s < pf; ++s) issue_load(s);
    }
    __syncwarp();
    double x = 0.0;  // x_i carried from the tile to the right
    for (int s = 0; s < nsteps; ++s) {
      if (elect_one()) {
        if (deep) bulk_wait_read<1>();  // the store of step s-2 has left its stage (= the stage of step s + ns - 2)
        else bulk_wait_read<0>();       // ... of step s-1 (= the stage of step s + ns - 1)
        issue_load(s + pf);
      }
      __syncwarp();
      const unsigned st = (it + (unsigned)s) % (unsigned)ns;
      mbar_wait(&full[st], ((it + (unsigned)s) / (unsigned)ns) & 1u, const_cast<int32_t *>(status));
      const int pan = b - s;
      unsigned char *stg = mystage + (size_t)st * STAGE;
      T *tile = reinterpret_cast<T *>(stg) + lane;  // element (lane, k) at tile[k * 32]
      const double *cf = reinterpret_cast<const double *>(stg + TILE);
      const bool diag = (s == 0);
      double lnew[kPanel];
#pragma unroll
      for (int k = kPanel - 1; k >= 0; --k) {
        const double lo = (double)tile[k * kPanel];
        const double cc = cf[k], sn = cf[kPanel + k];
        const double ln = cf[2 * kPanel + k] * (cc * lo - sn * x);
        const double xn = fma(cc, x, sn * lo);
        // above the diagonal (diagonal tile, k > lane): identity, the entry's bits never reach x or L
        const bool on = !diag || k <= lane;
        lnew[k] = ln;
        x = on ? xn : x;
      }
      if (!diag) {
#pragma unroll
        for (int k = 0; k < kPanel; ++k) tile[k * kPanel] = (T)lnew[k];
      } else if (valid) {
        T *gp = L + row + (int64_t)row0 * ld;
#pragma unroll
        for (int k = 0; k < kPanel; ++k)
          if (k <= lane) gp[(int64_t)k * ld] = (T)lnew[k];
      }
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        if (!diag) tma_store_2d(&tmapc, (int)row0, pan * kPanel, stg);
        bulk_commit();
      }
      __syncwarp();
    }
    it += (unsigned)nsteps;
  }
  if (elect_one()) bulk_wait_all<0>();
}

template <typename T>
inline size_t dd_sweep_cm2_smem() { return 1024 + (size_t)kSwCm2StagesTotal * SwCmGeom<T>::STAGE; }

// ---- third generation: FOUR warps per block row, each on 8 of a tile's 32 columns.  What bounds the kernels above
// is instruction issue, not bandwidth and not the dependent FMA chain: ~500 warp instructions per tile on ONE warp,
// 1.25 warps per scheduler, one instruction every 7.8 cycles (ncu, gpurun r04n) -- and the longest block row (T
// tiles in sequence) sets the kernel's time.  (Tried on the way: two warps taking alternate tiles with x handed over
// through shared memory, 0.82 ms; that plus packed coefficients, 0.84; that plus four-segment affine maps inside one
// warp, 0.95 -- more instructions, slower.)  The recurrence is affine in x, so warp h turns its 8 columns into
// (A_h, B_h): x behind the segment = A_h x_front + B_h, all four at once; the maps cross through shared memory (one
// named barrier of 128 threads), every warp composes the row's x across the tile for itself (4 FMAs) and the x in
// front of its own segment, then forms its 8 columns of L'.  A second barrier before the quad's TMA store.
constexpr int kSwCm4Quads = 4;   // block rows in flight per CTA (16 warps)
constexpr int kSwCm4Stages = 5;  // ring per block row; loads run three steps ahead

template <typename T>
__global__ void __launch_bounds__(kSwCm4Quads * 128, 1)
dd_sweep_cm4_kernel(T *L, int n, int64_t ld, LargeWs ws, const int32_t *status, const __grid_constant__ CUtensorMap tmapc) {
  constexpr int TILE = CmGeom<T>::TILE, STAGE = SwCmGeom<T>::STAGE, NS = kSwCm4Stages;
  extern __shared__ __align__(1024) unsigned char sw_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int q = warp >> 2, h = warp & 3;
  const int T_ = (n + 31) / 32;
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sw_smem) + q * NS;
  double2 *xch = reinterpret_cast<double2 *>(sw_smem + 1024) + (size_t)q * (2 * 4 * 32);  // [parity][segment][lane] (A, B)
  unsigned char *mystage = sw_smem + 1024 + kSwCm4Quads * 4096 + (size_t)q * NS * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSwCm4Quads * NS; ++i) mbar_init(reinterpret_cast<unsigned long long *>(sw_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  unsigned it = 0;  // tiles this quad has consumed so far (ring position / mbarrier phase / exchange parity)
  const int R = (T_ + G - 1) / G;  // rounds
  for (int j = q; j < R; j += kSwCm4Quads) {
    const int pos = j * G + ((j & 1) ? G - 1 - c : c);  // position in the list of rows sorted by length (snake)
    if (pos >= T_) continue;  // (last round only; the whole quad skips it)
    const int b = T_ - 1 - pos;
    const int64_t row0 = (int64_t)b * kPanel, row = row0 + lane;
    const bool valid = row < n;
    const int nsteps = b + 1;  // tiles b, b-1, ..., 0

    auto issue_load = [&](int s) {  // elected lane of warp h = 0
      if (s >= nsteps) return;
      const int pan = b - s;
      const unsigned st = (it + (unsigned)s) % NS;
      void *bar = &full[st];
      unsigned char *stg = mystage + (size_t)st * STAGE;
      mbar_arrive_expect_tx(bar, (unsigned)TILE + 1024u);
      tma_load_2d(stg, &tmapc, (int)row0, pan * kPanel, bar);
      bulk_g2s(stg + TILE, ws.X + (size_t)pan * (4 * kPanel), 1024u, bar);  // (c, s, sgn c, sgn s) per column
    };

    if (h == 0 && elect_one()) {
      bulk_wait_read<0>();  // (the previous row's last stores have left their stages)
      for (int s = 0; s < NS - 2; ++s) issue_load(s);
    }
    __syncwarp();
    double x = 0.0;  // x_i carried from the tile to the right (every warp of the quad keeps its own copy)
    for (int s = 0; s < nsteps; ++s) {
      if (h == 0 && elect_one()) {
        bulk_wait_read<1>();  // the store of step s-2 has left its stage (= the stage of step s + NS - 2)
        issue_load(s + NS - 2);
      }
      __syncwarp();
      const unsigned st = (it + (unsigned)s) % NS;
      mbar_wait(&full[st], ((it + (unsigned)s) / NS) & 1u, const_cast<int32_t *>(status));
      const int pan = b - s;
      unsigned char *stg = mystage + (size_t)st * STAGE;
      T *tile = reinterpret_cast<T *>(stg) + lane + (8 * h) * kPanel;  // element (lane, 8h + k) at tile[k * 32]
      const double2 *cf = reinterpret_cast<const double2 *>(stg + TILE) + 2 * (8 * h);  // column 8h + k: cf[2k] = (c, s), cf[2k+1] = sgn (c, s)
      const bool diag = (s == 0);
      double lo[8], sl[8], ck[8];
      double a = 1.0, bb = 0.0;
#pragma unroll
      for (int k = 7; k >= 0; --k) {
        const double2 cs = cf[2 * k];
        lo[k] = (double)tile[k * kPanel];
        ck[k] = cs.x;
        sl[k] = cs.y * lo[k];
        if (!diag || 8 * h + k <= lane) {  // above the diagonal (diagonal tile, column > lane): identity
          bb = fma(ck[k], bb, sl[k]);
          a *= ck[k];
        }
      }
      double2 *xs = xch + ((it + (unsigned)s) & 1u) * (4 * 32) + lane;  // segment g's map of this row at xs[g * 32]
      xs[h * 32] = make_double2(a, bb);
      named_sync(1 + q, 128);
      // x in front of my segment (segments 3, 2, ... are to my right), and x behind the whole tile
      double xf = x;
#pragma unroll
      for (int g = 3; g >= 0; --g) {
        const double2 m = xs[g * 32];
        x = fma(m.x, x, m.y);
        if (g == h + 1) xf = x;
      }
      T *gp = L + row + (int64_t)(row0 + 8 * h) * ld;  // (diagonal tile: straight to global memory, nothing above the diagonal)
#pragma unroll
      for (int k = 7; k >= 0; --k) {
        const double2 d = cf[2 * k + 1];
        const double ln = fma(-d.y, xf, d.x * lo[k]);
        if (!diag) {
          tile[k * kPanel] = (T)ln;
          xf = fma(ck[k], xf, sl[k]);
        } else if (8 * h + k <= lane) {
          if (valid) gp[(int64_t)k * ld] = (T)ln;
          xf = fma(ck[k], xf, sl[k]);
        }
      }
      if (!diag) {
        fence_proxy_async();
        named_sync(1 + q, 128);  // the other warps' tile writes too
        if (h == 0 && elect_one()) {
          tma_store_2d(&tmapc, (int)row0, pan * kPanel, stg);
          bulk_commit();
        }
      } else if (h == 0 && elect_one()) {
        bulk_commit();  // (keeps the wait_group arithmetic: one group per step)
      }
      __syncwarp();
    }
    it += (unsigned)nsteps;
  }
  if (h == 0 && elect_one()) bulk_wait_all<0>();
}

template <typename T>
inline size_t dd_sweep_cm4_smem() { return 1024 + kSwCm4Quads * 4096 + (size_t)kSwCm4Quads * kSwCm4Stages * SwCmGeom<T>::STAGE; }

template <typename T>
inline size_t dd_sweep_cm_smem();

template <typename T>
inline cudaError_t launch_dd_sweep_cm(T *L, int n, int64_t ld, const LargeWs &ws, const int32_t *status,
                                      const CUtensorMap &tmapc, cudaStream_t st) {
  const int nb = (n + 31) / 32;
  if (getenv("CHUB_DD_SWEEP_CM_V1")) {  // A/B: four consecutive block rows per CTA
    auto sk = dd_sweep_cm_kernel<T>;
    const size_t ssm = dd_sweep_cm_smem<T>();
    cudaError_t e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
    if (e != cudaSuccess) return e;
    sk<<<(nb + 3) / 4, 128, ssm, st>>>(L, n, ld, ws, status, tmapc);
    return cudaSuccess;
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!getenv("CHUB_DD_SWEEP_CM_V2") && !getenv("CHUB_DD_COEF_V1")) {  // (V2: one warp per block row, A/B; the packed
    // coefficients this kernel reads come from dd_coef_par_kernel only)
    auto sk = dd_sweep_cm4_kernel<T>;
    const size_t ssm = dd_sweep_cm4_smem<T>();
    cudaError_t e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
    if (e != cudaSuccess) return e;
    sk<<<std::max(1, std::min(sms, nb)), kSwCm4Quads * 128, ssm, st>>>(L, n, ld, ws, status, tmapc);
    return cudaSuccess;
  }
  auto sk = dd_sweep_cm2_kernel<T>;
  const size_t ssm = dd_sweep_cm2_smem<T>();
  cudaError_t e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
  if (e != cudaSuccess) return e;
  const int grid = std::max(1, std::min(sms, nb));
  const int rounds = (nb + grid - 1) / grid;  // block rows per CTA
  int ns = std::max(3, std::min(8, kSwCm2StagesTotal / std::min(kSwCm2Warps, rounds)));
  if (const char *e = getenv("CHUB_DD_SWEEP_CM_NS")) ns = std::max(3, std::min(ns, atoi(e)));
  sk<<<grid, kSwCm2Warps * 32, ssm, st>>>(L, n, ld, ws, status, ns, tmapc);
  return cudaSuccess;
}

template <typename T>
inline size_t dd_sweep_cm_smem() { return 1024 + (size_t)4 * kSwCmStages * SwCmGeom<T>::STAGE; }

// ---- rectangle below a diagonal block, column-major (blocked driver for factors beyond the persistent kernel's
// row capacity; the column-major twin of rect_apply_rm_kernel).  A warp takes a block row of 32 rows (lane = row)
// and walks the diagonal block's columns in [32 rows x 32 columns] TMA boxes through a ring, the block column's
// (p, c, sigma) riding along; nothing here waits for anybody, so the row's residual chain stays a plain
// sequential FMA chain per lane (six warps per SM hide it: the kernel is a pure stream).
//     L_ik <- c_k L_ik + sigma_k w_i ,   w_i <- w_i - L_ik p_k
// UPDATE = false: residual only (forward substitution of the downdate).
constexpr int kRectCmStages = 4;
constexpr int kRectCmWarps = 6;

template <typename T>
struct RectCmGeom {
  static constexpr int STAGE = CmGeom<T>::TILE + 1024;  // tile + (p, c, sigma)[32] doubles (768 B, padded)
};

template <typename T, bool UPDATE>
__global__ void __launch_bounds__(kRectCmWarps * 32, 1)
rect_apply_cm_kernel(int n, int row_begin, int col_begin, int ncols, LargeWs ws, const int32_t *status,
                     const __grid_constant__ CUtensorMap tmapc) {
  constexpr int TILE = CmGeom<T>::TILE, STAGE = RectCmGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char rcm_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(rcm_smem) + warp * kRectCmStages;
  unsigned char *mystage = rcm_smem + 1024 + (size_t)warp * kRectCmStages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRectCmWarps * kRectCmStages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(rcm_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int nsteps = ncols / kPanel;
  const int ngroups = (n - row_begin + kPanel - 1) / kPanel;
  unsigned it = 0;  // tiles this warp has consumed so far (ring position / mbarrier phase)
  // (groups dealt CTA-first: with fewer groups than warp slots every SM still gets its share)
  for (int grp = (int)blockIdx.x + warp * (int)gridDim.x; grp < ngroups; grp += (int)gridDim.x * kRectCmWarps) {
    const int row0 = row_begin + grp * kPanel;
    const int row = row0 + lane;
    double w = row < n ? ws.w[row] : 0.0;

    auto issue_load = [&](int s) {  // elected lane; stage / phase follow the warp's running tile count
      if (s >= nsteps) return;
      const unsigned st = (it + (unsigned)s) % kRectCmStages;
      void *bar = &full[st];
      unsigned char *stg = mystage + (size_t)st * STAGE;
      const int c0 = col_begin + s * kPanel;
      mbar_arrive_expect_tx(bar, (unsigned)TILE + (UPDATE ? 768u : 256u));
      tma_load_2d(stg, &tmapc, row0, c0, bar);
      bulk_g2s(stg + TILE, ws.p + c0, 256u, bar);
      if (UPDATE) {
        bulk_g2s(stg + TILE + 256, ws.c + c0, 256u, bar);
        bulk_g2s(stg + TILE + 512, ws.sg + c0, 256u, bar);
      }
    };

    if (elect_one()) {
      bulk_wait_read<0>();  // (the previous group's last stores have left their stages)
      for (int s = 0; s < kRectCmStages - 1; ++s) issue_load(s);
    }
    __syncwarp();
    for (int s = 0; s < nsteps; ++s) {
      const unsigned st = (it + (unsigned)s) % kRectCmStages;
      const unsigned ph = ((it + (unsigned)s) / kRectCmStages) & 1u;
      mbar_wait(&full[st], ph, const_cast<int32_t *>(status));
      unsigned char *stg = mystage + (size_t)st * STAGE;
      T *tile = reinterpret_cast<T *>(stg) + lane;  // element (lane, k) at tile[k * 32]
      const double *cf = reinterpret_cast<const double *>(stg + TILE);
#pragma unroll
      for (int k = 0; k < kPanel; k += 2) {
        const double2 pp = *reinterpret_cast<const double2 *>(cf + k);
        const double l0 = (double)tile[k * kPanel], l1 = (double)tile[(k + 1) * kPanel];
        if (UPDATE) {
          const double2 c2 = *reinterpret_cast<const double2 *>(cf + kPanel + k);
          const double2 s2 = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + k);
          tile[k * kPanel] = (T)fma(c2.x, l0, s2.x * w);
          w = fma(-l0, pp.x, w);
          tile[(k + 1) * kPanel] = (T)fma(c2.y, l1, s2.y * w);
          w = fma(-l1, pp.y, w);
        } else {
          w = fma(-l0, pp.x, w);
          w = fma(-l1, pp.y, w);
        }
      }
      if (UPDATE) fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        if (UPDATE) {
          tma_store_2d(&tmapc, row0, col_begin + s * kPanel, stg);
          bulk_commit();
          bulk_wait_read<1>();  // the store of step s-1 has left the stage that step s+3 loads into
        }
        issue_load(s + kRectCmStages - 1);
      }
      __syncwarp();
    }
    it += (unsigned)nsteps;
    if (row < n) ws.w[row] = w;
  }
  if (elect_one()) bulk_wait_all<0>();
}

template <typename T>
inline size_t rect_apply_cm_smem() { return 1024 + (size_t)kRectCmWarps * kRectCmStages * RectCmGeom<T>::STAGE; }

// =====================================================================================
// Downdate sweep (row-major): TMA-pipelined, 4 lanes per row, right-to-left affine scan
// =====================================================================================
// Reverse sweep of _seeger_impl_cython.pyx:285-330 in the row-independent form of SURVEY A.3:
//   (L_ik, x_i) <- (sgn_k (c_k L_ik - s_k x_i), c_k x_i + s_k L_ik)   for k = i .. 0,  x_i = 0 at k = i.
// Rows are independent and all (c_k, s_k, sgn_k) are known, so there is no chain: a CTA takes one
// block row (4 warps x 8 rows) and every warp walks its tiles from the diagonal to column 0 with a
// ring of TMA loads one step ahead.  x -> c x + s L is affine in x, so lane (g, r8) first composes
// the map of its 8 columns, a two-step shuffle scan (right to left over the 4 lanes of a row) gives
// every lane its incoming x, and a second local pass writes the new entries.
constexpr int kSwStages = 4;

template <typename T>
struct SwGeom {
  static constexpr int STAGE = RmGeom<T>::WTILE + 1024;  // tile + (c, s, sgn)[32] doubles (768 B, padded)
};

template <typename T>
__global__ void __launch_bounds__(128) dd_sweep_rm_kernel(T *L, int n, int64_t ld, LargeWs ws,
                                                          const int32_t *status,
                                                          const __grid_constant__ CUtensorMap tmap8) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, CPT = RmGeom<T>::CPT;
  constexpr int WTILE = RmGeom<T>::WTILE, STAGE = SwGeom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char sw_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int T_ = (n + 31) / 32;
  const int b = T_ - 1 - blockIdx.x;  // longest rows first
  const int64_t row0 = (int64_t)b * 32 + warp * 8;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sw_smem) + warp * kSwStages;
  unsigned char *mystage = sw_smem + 1024 + (size_t)warp * kSwStages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * kSwStages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(sw_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (row0 >= n) return;
  const int g = lane >> 3, r8 = lane & 7;
  const int64_t row = row0 + r8;
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  const unsigned my_off = hh * 1024 + r8 * 128;
  const int nsteps = b + 1;  // tiles b, b-1, ..., 0

  auto issue_load = [&](int s, int st) {  // elected lane
    if (s >= nsteps) return;
    const int pan = b - s;
    void *bar = &full[st];
    unsigned char *stg = mystage + (size_t)st * STAGE;
    const int c0 = pan * kPanel;
    const bool two = HALVES == 2 && c0 + BOXC < n;
    mbar_arrive_expect_tx(bar, (two ? 2048u : 1024u) + 768u);
    tma_load_2d(stg, &tmap8, c0, (int)row0, bar);
    if (two) tma_load_2d(stg + 1024, &tmap8, c0 + BOXC, (int)row0, bar);
    bulk_g2s(stg + WTILE, ws.c + c0, 256u, bar);
    bulk_g2s(stg + WTILE + 256, ws.sg + c0, 256u, bar);
    bulk_g2s(stg + WTILE + 512, ws.dn + c0, 256u, bar);
  };

  if (elect_one())
    for (int s = 0; s < kSwStages - 2; ++s) issue_load(s, s);
  __syncwarp();

  double xc = 0.0;  // x_i carried from the tile to the right
  int st = 0, stn = kSwStages - 2;
  unsigned ph = 0;
  for (int s = 0; s < nsteps; ++s) {
    if (elect_one()) {
      bulk_wait_read<1>();
      issue_load(s + kSwStages - 2, stn);
    }
    __syncwarp();
    mbar_wait(&full[st], ph, const_cast<int32_t *>(status));
    const int pan = b - s;
    unsigned char *stg = mystage + (size_t)st * STAGE;
    unsigned char *rowp = stg + my_off;
    const double *cf = reinterpret_cast<const double *>(stg + WTILE) + g * 8;
    const int64_t c0 = (int64_t)pan * kPanel + g * 8;
    const bool diag_step = (s == 0);
    int4 raw[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) raw[i] = *reinterpret_cast<const int4 *>(rowp + (((cb + i) ^ r8) << 4));
    T *x = reinterpret_cast<T *>(raw);
    double cc[8], sc[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 a = *reinterpret_cast<const double2 *>(cf + j);
      const double2 d = *reinterpret_cast<const double2 *>(cf + kPanel + j);
      cc[j] = a.x; cc[j + 1] = a.y; sc[j] = d.x; sc[j + 1] = d.y;
    }
    // pass 1: the affine map x -> A x + B of my 8 columns (right to left)
    double A = 1.0, B = 0.0;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      lo[j] = (double)x[j];
      if (!diag_step || c0 + j <= row) {
        B = fma(cc[j], B, sc[j] * lo[j]);
        A *= cc[j];
      } else {
        // above the diagonal: identity map, entry left alone; its bits (never read by the reference,
        // possibly NaN) must not reach the arithmetic
        cc[j] = 1.0; sc[j] = 0.0; lo[j] = 0.0;
      }
    }
    // inclusive suffix composition over the 4 lanes of the row: P_g = M_g o M_{g+1} o ... o M_3
    double PA = A, PB = B;
    double qa = __shfl_down_sync(0xffffffffu, PA, 8), qb = __shfl_down_sync(0xffffffffu, PB, 8);
    if (g <= 2) { PB = fma(PA, qb, PB); PA *= qa; }
    qa = __shfl_down_sync(0xffffffffu, PA, 16); qb = __shfl_down_sync(0xffffffffu, PB, 16);
    if (g <= 1) { PB = fma(PA, qb, PB); PA *= qa; }
    // incoming x of my segment = (M_{g+1} o ... o M_3)(xc);  outgoing x of the tile = P_0(xc)
    const double ea = __shfl_down_sync(0xffffffffu, PA, 8), eb = __shfl_down_sync(0xffffffffu, PB, 8);
    double xi = g == 3 ? xc : fma(ea, xc, eb);
    const double xout = fma(__shfl_sync(0xffffffffu, PA, r8), xc, __shfl_sync(0xffffffffu, PB, r8));
    // pass 2: new entries
    const double *sgp = cf + 2 * kPanel;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      const double lnew = sgp[j] * (cc[j] * lo[j] - sc[j] * xi);
      xi = fma(cc[j], xi, sc[j] * lo[j]);
      if (!diag_step || c0 + j <= row) x[j] = (T)lnew;
    }
    xc = xout;
    if (!diag_step) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
    } else if (row < n) {
      T *gp = L + row * ld + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j <= row) gp[j] = x[j];
    }
    fence_proxy_async();
    __syncwarp();
    if (elect_one()) {
      if (!diag_step) {
        const int cc0 = pan * kPanel;
        tma_store_2d(&tmap8, cc0, (int)row0, stg);
        if (HALVES == 2 && cc0 + BOXC < n) tma_store_2d(&tmap8, cc0 + BOXC, (int)row0, stg + 1024);
      }
      bulk_commit();
    }
    __syncwarp();
    if (++st == kSwStages) { st = 0; ph ^= 1u; }
    if (++stn == kSwStages) stn = 0;
  }
  if (elect_one()) bulk_wait_all<0>();
}

template <typename T>
inline size_t dd_sweep_rm_smem() { return 1024 + (size_t)4 * kSwStages * SwGeom<T>::STAGE; }

// ---- downdate sweep, row-major, second generation: after the (ragged) diagonal tile a warp walks its rows right to
// left in ITERATIONS of two block columns -- one 3-D tensor-map TMA load and one store per [8 rows][64 columns] tile
// and three 512-byte bulk copies for (c, s, sgn) instead of fourteen TMA operations per 64 columns (the
// first-generation kernel above was bound by the issue of those: 480 us = 68 % of the HBM roofline at n = 16384,
// profiles/r03n).  Same arithmetic and lane mapping.
constexpr int kSw2Stages = 3;

template <typename T>
struct Sw2Geom {
  static constexpr int COEF = 3 * 2 * kPanel * 8;                                  // (c, s, sgn)[64] doubles
  static constexpr int STAGE = ((2 * RmGeom<T>::WTILE + COEF + 1023) / 1024) * 1024;  // tile pair + coefficients
};

template <typename T>
__global__ void __launch_bounds__(128) dd_sweep_rm2_kernel(T *L, int n, int64_t ld, LargeWs ws,
                                                           const int32_t *status, int use3d,
                                                           const __grid_constant__ CUtensorMap tmap8,
                                                           const __grid_constant__ CUtensorMap tmap3) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, CPT = RmGeom<T>::CPT;
  constexpr int WTILE = RmGeom<T>::WTILE, STAGE = Sw2Geom<T>::STAGE;
  extern __shared__ __align__(1024) unsigned char sw_smem[];
  if (*status != 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int T_ = (n + 31) / 32;
  const int b = T_ - 1 - blockIdx.x;  // longest rows first
  const int64_t row0 = (int64_t)b * 32 + warp * 8;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sw_smem) + warp * kSw2Stages;
  unsigned char *mystage = sw_smem + 1024 + (size_t)warp * kSw2Stages * STAGE;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * kSw2Stages; ++i) mbar_init(reinterpret_cast<unsigned long long *>(sw_smem) + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (row0 >= n) return;
  const int g = lane >> 3, r8 = lane & 7;
  const int64_t row = row0 + r8;
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  const unsigned my_off = hh * 1024 + r8 * 128;
  const int n3 = use3d ? (n / BOXC) * BOXC : 0;
  // iterations: 0 = the diagonal tile b; then pairs (b-1, b-2), (b-3, b-4), ...; a last single tile 0 if b is odd
  const int NI = 1 + (b + 1) / 2;
  auto iter = [&](int it, int &pan_lo, int &cnt) {  // block columns pan_lo .. pan_lo + cnt - 1
    if (it == 0) { pan_lo = b; cnt = 1; return; }
    const int hi = b - 1 - 2 * (it - 1);
    if (hi >= 1) { pan_lo = hi - 1; cnt = 2; } else { pan_lo = 0; cnt = 1; }
  };

  auto issue_load = [&](int it) {  // elected lane
    if (it >= NI) return;
    int pan_lo, cnt;
    iter(it, pan_lo, cnt);
    const int st = it % kSw2Stages;
    void *bar = &full[st];
    unsigned char *stg = mystage + (size_t)st * STAGE;
    const int c0 = pan_lo * kPanel;
    const unsigned cbytes = (unsigned)cnt * kPanel * 8u;
    if (cnt == 2 && c0 + 2 * kPanel <= n3) {
      mbar_arrive_expect_tx(bar, 2u * WTILE + 3u * cbytes);
      tma_load_3d(stg, &tmap3, 0, (int)row0, c0 / BOXC, bar);
    } else {
      unsigned bytes = 0;
      for (int t = 0; t < cnt; ++t) bytes += (HALVES == 2 && c0 + t * kPanel + BOXC < n) ? 2048u : 1024u;
      mbar_arrive_expect_tx(bar, bytes + 3u * cbytes);
      for (int t = 0; t < cnt; ++t) {
        const int cc = c0 + t * kPanel;
        tma_load_2d(stg + t * WTILE, &tmap8, cc, (int)row0, bar);
        if (HALVES == 2 && cc + BOXC < n) tma_load_2d(stg + t * WTILE + 1024, &tmap8, cc + BOXC, (int)row0, bar);
      }
    }
    // coefficient area: [c | s | sgn], each 64 doubles (the first cnt * 32 used), indexed by column - c0
    bulk_g2s(stg + 2 * WTILE, ws.c + c0, cbytes, bar);
    bulk_g2s(stg + 2 * WTILE + 512, ws.sg + c0, cbytes, bar);
    bulk_g2s(stg + 2 * WTILE + 1024, ws.dn + c0, cbytes, bar);
  };

  if (elect_one()) {
    issue_load(0);
    if (kSw2Stages > 2) issue_load(1);
  }
  __syncwarp();

  double xc = 0.0;  // x_i carried from the tile to the right
  // one tile: tile bytes at tp, coefficients of its 32 columns at cf0 (c), cf0 + 64 (s), cf0 + 128 (sgn)
  auto do_tile = [&](unsigned char *tp, const double *cf0, int pan, bool diag_step) {
    unsigned char *rowp = tp + my_off;
    const double *cf = cf0 + g * 8;
    const int64_t c0 = (int64_t)pan * kPanel + g * 8;
    int4 raw[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) raw[i] = *reinterpret_cast<const int4 *>(rowp + (((cb + i) ^ r8) << 4));
    T *x = reinterpret_cast<T *>(raw);
    double cc[8], sc[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 a = *reinterpret_cast<const double2 *>(cf + j);
      const double2 d = *reinterpret_cast<const double2 *>(cf + 64 + j);
      cc[j] = a.x; cc[j + 1] = a.y; sc[j] = d.x; sc[j + 1] = d.y;
    }
    // pass 1: the affine map x -> A x + B of my 8 columns (right to left)
    double A = 1.0, B = 0.0;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      lo[j] = (double)x[j];
      if (!diag_step || c0 + j <= row) {
        B = fma(cc[j], B, sc[j] * lo[j]);
        A *= cc[j];
      } else {
        // above the diagonal: identity map, entry left alone; its bits (never read by the reference,
        // possibly NaN) must not reach the arithmetic
        cc[j] = 1.0; sc[j] = 0.0; lo[j] = 0.0;
      }
    }
    // inclusive suffix composition over the 4 lanes of the row: P_g = M_g o M_{g+1} o ... o M_3
    double PA = A, PB = B;
    double qa = __shfl_down_sync(0xffffffffu, PA, 8), qb = __shfl_down_sync(0xffffffffu, PB, 8);
    if (g <= 2) { PB = fma(PA, qb, PB); PA *= qa; }
    qa = __shfl_down_sync(0xffffffffu, PA, 16); qb = __shfl_down_sync(0xffffffffu, PB, 16);
    if (g <= 1) { PB = fma(PA, qb, PB); PA *= qa; }
    // incoming x of my segment = (M_{g+1} o ... o M_3)(xc);  outgoing x of the tile = P_0(xc)
    const double ea = __shfl_down_sync(0xffffffffu, PA, 8), eb = __shfl_down_sync(0xffffffffu, PB, 8);
    double xi = g == 3 ? xc : fma(ea, xc, eb);
    const double xout = fma(__shfl_sync(0xffffffffu, PA, r8), xc, __shfl_sync(0xffffffffu, PB, r8));
    // pass 2: new entries
    const double *sgp = cf + 128;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      const double lnew = sgp[j] * (cc[j] * lo[j] - sc[j] * xi);
      xi = fma(cc[j], xi, sc[j] * lo[j]);
      if (!diag_step || c0 + j <= row) x[j] = (T)lnew;
    }
    xc = xout;
    if (!diag_step) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
    } else if (row < n) {
      T *gp = L + row * ld + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j <= row) gp[j] = x[j];
    }
  };

  for (int it = 0; it < NI; ++it) {
    int pan_lo, cnt;
    iter(it, pan_lo, cnt);
    const int st = it % kSw2Stages;
    mbar_wait(&full[st], (unsigned)((it / kSw2Stages) & 1), const_cast<int32_t *>(status));
    unsigned char *stg = mystage + (size_t)st * STAGE;
    const double *cfb = reinterpret_cast<const double *>(stg + 2 * WTILE);
    const bool diag_step = it == 0;
    if (cnt == 2) do_tile(stg + WTILE, cfb + kPanel, pan_lo + 1, false);  // right tile first (right-to-left sweep)
    do_tile(stg, cfb, pan_lo, diag_step);
    // loads run kSw2Stages - 1 iterations ahead; iteration it + stages - 1 reuses the stage of iteration it - 1,
    // whose store has had this compute phase to leave shared memory
    if (elect_one()) {
      if (it >= 1) bulk_wait_read<0>();
      issue_load(it + kSw2Stages - 1);
    }
    fence_proxy_async();
    __syncwarp();
    if (elect_one()) {
      if (!diag_step) {
        const int c0 = pan_lo * kPanel;
        if (cnt == 2 && c0 + 2 * kPanel <= n3) {
          tma_store_3d(&tmap3, 0, (int)row0, c0 / BOXC, stg);
        } else {
          for (int t = 0; t < cnt; ++t) {
            const int cc = c0 + t * kPanel;
            tma_store_2d(&tmap8, cc, (int)row0, stg + t * WTILE);
            if (HALVES == 2 && cc + BOXC < n) tma_store_2d(&tmap8, cc + BOXC, (int)row0, stg + t * WTILE + 1024);
          }
        }
      }
      bulk_commit();
    }
    __syncwarp();
  }
  if (elect_one()) bulk_wait_all<0>();
}

template <typename T>
inline size_t dd_sweep_rm2_smem() { return 1024 + (size_t)4 * kSw2Stages * Sw2Geom<T>::STAGE; }

// ------------------------------------------------------------------ host side
inline size_t dataflow_flag_bytes(int64_t n) { return 256 + (256 * 16 + 8 * ((size_t)n / 32 + 2)) * sizeof(long long); }

struct DfPlan {
  bool ok = false;
  int sms = 0, R = 0, r_shift = 0, Wk = 0, rows_max = 0, nw = 0;
  size_t smem = 0;
};

template <typename T>
inline DfPlan df_plan(int64_t n, int64_t ld, int layout, const void *L) {
  DfPlan pl;
  constexpr int EPC = 16 / sizeof(T);
  if (n < 64 || n > (1 << 30) || (ld % EPC) != 0 || (reinterpret_cast<uintptr_t>(L) & 15) != 0) return pl;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pl;
  int sms = 0, coop = 0, smem_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (!coop || sms < 2) return pl;
  const int Wk = sms - 1;
  int best_R = 0, best_rows = 1 << 30;
  // rows are dealt in blocks of R; column-major needs R = 16 (one 16-row block per half warp)
  for (int R = 16; R >= (layout == CHUB_COL_MAJOR ? 16 : 1); R >>= 1) {
    const int64_t nRB = (n + R - 1) / R;
    const int rows = (int)((nRB + Wk - 1) / Wk) * R;
    if (rows < best_rows) { best_rows = rows; best_R = R; }
  }
  const int nw = (best_rows + 31) / 32;
  if (best_R == 0 || nw > kDfRowWarps) return pl;
  const size_t smem = std::max(df_chain_smem<T>(layout), df_worker_smem<T>(layout, nw));
  if (smem > (size_t)smem_optin) return pl;
  pl.ok = true; pl.sms = sms; pl.R = best_R; pl.Wk = Wk; pl.rows_max = best_rows; pl.nw = nw; pl.smem = smem;
  while ((1 << pl.r_shift) < best_R) ++pl.r_shift;
  return pl;
}

// ---- v3 plan (row-major): rows dealt in blocks of 16, two warps per block
struct RmPlan {
  bool ok = false;
  int sms = 0, Wk = 0, Q = 0, nrw = 0, threads = 0;
  size_t smem = 0;
};

template <typename T>
inline RmPlan rm_plan(int64_t n, int64_t ld, const void *L) {
  RmPlan pl;
  constexpr int EPC = 16 / sizeof(T);
  if (n < 64 || n >= (1LL << 31) || (ld % EPC) != 0 || (reinterpret_cast<uintptr_t>(L) & 15) != 0) return pl;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pl;
  int sms = 0, coop = 0, smem_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (!coop || sms < 2) return pl;
  const int Wk = sms - 1;
  const int64_t nRB = (n + 15) / 16;
  const int Q = (int)((nRB + Wk - 1) / Wk);
  if (Q > kRmMaxQ) return pl;
  const int nrw = 2 * Q;
  size_t smem = std::max(df_chain_rm_smem<T>(), df_worker_rm_smem<T>(nrw));
#ifdef CHUB_CHAIN_V4
  smem = std::max(smem, df_chain_v4_smem<T>());
#endif
  smem = std::max(smem, df_chain_v6_smem<T>());
#if CHUB_CHAIN_V7
  smem = std::max(smem, df_chain_v7_smem<T>());
#endif
  if (smem > (size_t)smem_optin) return pl;
  pl.ok = true; pl.sms = sms; pl.Wk = Wk; pl.Q = Q; pl.nrw = nrw;
  pl.threads = std::max(std::max(256, kC6ThreadsSplit), 32 * (nrw + 2));  // row warps + coefficient warp + relay warp
  pl.smem = smem;
  return pl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// Row-major n x n matrix with leading dimension ld: boxes of [box_rows][128 bytes], SWIZZLE_128B.
template <typename T>
inline bool make_tmap(CUtensorMap *map, T *L, int64_t n, int64_t ld, int box_rows, int64_t rows = -1) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)(rows < 0 ? n : rows)};  // (columns, rows)
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)RmGeom<T>::BOXC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  return enc(map, dt, 2, L, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The same matrix as [column block of BOXC][row][BOXC columns]: a box of 2*HALVES column blocks x 8 rows is a
// [8 rows][64 columns] tile in ONE TMA operation, landing as the 128-byte-swizzled 1 KB boxes make_tmap gives.
// (Only whole column blocks are visible: the ragged right edge of an n that is no multiple of BOXC goes
// through the 2-D map.)
template <typename T>
inline bool make_tmap3(CUtensorMap *map, T *L, int64_t n, int64_t ld, int box_rows, int box_cblocks) {
  EncodeTiledFn enc = get_encode_tiled();
  constexpr int BOXC = RmGeom<T>::BOXC;
  if (!enc || n / BOXC < box_cblocks) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)BOXC, (cuuint64_t)n, (cuuint64_t)(n / BOXC)};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(T), (cuuint64_t)BOXC * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)BOXC, (cuuint32_t)box_rows, (cuuint32_t)box_cblocks};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  return enc(map, dt, 3, L, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- column-major plan: block rows of 32 dealt cyclically, one warp per block row
struct CmPlan {
  bool ok = false;
  int sms = 0, Wk = 0, Q = 0;
  size_t smem = 0;
};

template <typename T>
inline CmPlan cm_plan(int64_t n, int64_t ld, const void *L) {
  CmPlan pl;
  constexpr int EPC = 16 / sizeof(T);
  if (n < 64 || n >= (1LL << 31) || (ld % EPC) != 0 || (reinterpret_cast<uintptr_t>(L) & 15) != 0) return pl;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pl;
  int sms = 0, coop = 0, smem_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (!coop || sms < 2) return pl;
  const int Wk = sms - 1;
  const int64_t nRB = (n + kPanel - 1) / kPanel;
  const int Q = (int)((nRB + Wk - 1) / Wk);
  if (Q > kCmMaxQ) return pl;
  size_t smem = std::max(std::max(df_chain_rm_smem<T>(), df_chain_v6_smem<T>()), df_worker_cm2_smem<T>(Q, Q <= kCm4MaxQ ? 4 : 2));
#if CHUB_CHAIN_V7
  smem = std::max(smem, df_chain_v7_smem<T>());
#endif
  if (smem > (size_t)smem_optin) return pl;
  pl.ok = true; pl.sms = sms; pl.Wk = Wk; pl.Q = Q; pl.smem = smem;
  return pl;
}

// Column-major n x n matrix with leading dimension ld: plain boxes of [32 columns][32 rows] (rows contiguous).
template <typename T>
inline bool make_tmap_cm(CUtensorMap *map, T *L, int64_t n, int64_t ld, int box_rows = kPanel) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)n};  // (rows, columns): rows are the inner dimension
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)kPanel};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  return enc(map, dt, 2, L, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The tensor-map kernels (row-major df_kernel_rm, column-major df_kernel_cm) take this shape natively.
template <typename T>
inline bool dataflow_native(int64_t n, int64_t ld, int layout, const T *L) {
  if (!get_encode_tiled() || getenv("CHUB_DF_V2")) return false;
  if (layout == CHUB_COL_MAJOR) {  // (test knob: send smaller column-major factors through the blocked driver)
    if (const char *e = getenv("CHUB_CM_NATIVE_MAX_N")) {
      if (n > atoll(e)) return false;
    }
  }
  return layout == CHUB_ROW_MAJOR ? rm_plan<T>(n, ld, L).ok : cm_plan<T>(n, ld, L).ok;
}

template <typename T>
inline bool dataflow_supported(int64_t n, int64_t ld, int layout, const T *L) {
  if (dataflow_native<T>(n, ld, layout, L)) return true;
  return df_plan<T>(n, ld, layout, L).ok;
}

// Launches prepare + diagonal inverses + the persistent kernel (+ the downdate's coefficient and
// sweep kernels).  Returns a cudaError_t-style code; *launches = kernels launched.
template <typename T, int LAYOUT, bool UPDATE>
int launch_dataflow(T *L, int64_t n, int64_t ld, T *v, int flags, const LargeWs &ws,
                    int32_t *status, cudaStream_t st, int64_t *launches, const double *w_in = nullptr,
                    bool chain_only = false) {
  // w_in / chain_only: blocked driver (launch_blocked) -- this call handles one diagonal block whose
  // input vector is the residual w_in, and for a downdate only the forward substitution (the
  // coefficient and sweep kernels run once, over the whole factor, afterwards)
  *launches = 0;
  const int ni = (int)n;
  const int n_pad = (ni + 31) & ~31;
  const int nb = n_pad / 32;
  RmPlan rp;
  alignas(64) CUtensorMap tmap8, tmap32, tmap3;
  bool use3d = false;
  if (LAYOUT == CHUB_ROW_MAJOR && !getenv("CHUB_DF_V2")) {
    rp = rm_plan<T>(n, ld, L);
    // tmap32: the chain's box (all d band tiles of a block column); tmap8 / tmap3: the row warps' boxes
#if CHUB_CHAIN_V7  // (the distance-1 and -2 tiles are not loaded: M1 and M2 stand for them)
    if (rp.ok && !(make_tmap<T>(&tmap8, L, n, ld, 8) &&
                   make_tmap3<T>(&tmap32, L, n, ld, kPanel * (kDfD - 2), RmGeom<T>::HALVES)))
      rp.ok = false;
#elif defined(CHUB_CHAIN_V4)
    if (rp.ok && !(make_tmap<T>(&tmap8, L, n, ld, 8) &&
                   make_tmap3<T>(&tmap32, L, n, ld, kPanel * (kDfD - 1), RmGeom<T>::HALVES)))
      rp.ok = false;
#else
    if (rp.ok && !(make_tmap<T>(&tmap8, L, n, ld, 8) &&
                   make_tmap3<T>(&tmap32, L, n, ld, kPanel * kDfD, RmGeom<T>::HALVES)))
      rp.ok = false;
#endif
    if (rp.ok) {
      const char *e3 = getenv("CHUB_RM5_3D");
      use3d = !(e3 && e3[0] == '0') && make_tmap3<T>(&tmap3, L, n, ld, 8, 2 * RmGeom<T>::HALVES);
      if (!use3d) tmap3 = tmap8;
    }
  }
  CmPlan cp;
  alignas(64) CUtensorMap tmapc, tmapcb;  // tmapcb: the chain's box (32 d rows x 32 columns)
  if (LAYOUT == CHUB_COL_MAJOR && !getenv("CHUB_DF_V2")) {
    cp = cm_plan<T>(n, ld, L);
    if (cp.ok && !(make_tmap_cm<T>(&tmapc, L, n, ld) && make_tmap_cm<T>(&tmapcb, L, n, ld, kPanel * (CHUB_CHAIN_V7 ? kDfD - 2 : kDfD))))
      cp.ok = false;
  }
  const bool tm = rp.ok || cp.ok;  // a tensor-map kernel (band depth kDfD)
  DfPlan pl;
  if (!tm) {
    pl = df_plan<T>(n, ld, LAYOUT, L);
    if (!pl.ok) return (int)cudaErrorNotSupported;
  }
  // one pre-pass: w-tilde / v0 / dg / sentinels (prepare) and the inverse diagonal blocks
  df_prepass_kernel<T, LAYOUT><<<(nb + 3) / 4, 128, 0, st>>>(L, ni, ld, v, ws, flags, status,
#if CHUB_CHAIN_V7
                                                              tm ? kDfD : kV2D, tm, w_in);
#elif defined(CHUB_CHAIN_V4)
                                                              tm ? kDfD : kV2D, rp.ok, w_in);
#else
                                                              tm ? kDfD : kV2D, false, w_in);
#endif
  ++*launches;
  DfParams P;
  P.n = ni; P.T = nb; P.ld = ld; P.ws = ws; P.status = status;
  P.prof = nullptr;
  if (const char *pe = getenv("CHUB_PROFILE")) {
    P.prof_late = pe[0] == '2';
    P.prof = reinterpret_cast<long long *>(reinterpret_cast<char *>(ws.flags) + 256);
    cudaMemsetAsync(P.prof, 0, (256 * 16 + 8 * ((size_t)n / 32 + 2)) * sizeof(long long), st);
  }
  cudaError_t e;
  if (rp.ok) {
    P.R = 16; P.r_shift = 4; P.Wk = rp.Wk; P.rows_max = rp.Q * 16; P.nw = rp.nrw;
    P.use3d = use3d ? 1 : 0;
    const char *e5 = getenv("CHUB_RM5");
    const bool v5 = !(e5 && e5[0] == '0');
    auto kern = v5 ? (P.prof ? df_kernel_rm<T, UPDATE, true, true> : df_kernel_rm<T, UPDATE, false, true>)
                   : (P.prof ? df_kernel_rm<T, UPDATE, true, false> : df_kernel_rm<T, UPDATE, false, false>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp.smem);
    if (e != cudaSuccess) return (int)e;
    void *args[] = {(void *)&L, (void *)&v, (void *)&P, (void *)&tmap8, (void *)&tmap32, (void *)&tmap3};
    e = cudaLaunchCooperativeKernel((void *)kern, dim3(rp.sms), dim3(rp.threads), args, rp.smem, st);
  } else if (cp.ok) {
    P.R = 32; P.r_shift = 5; P.Wk = cp.Wk; P.rows_max = cp.Q * 32; P.nw = cp.Q;
    const char *ep = getenv("CHUB_CM_PAIR");
    // warps per block row: 4 where the thread count allows, else 2 (df_worker_cm2); 0 / 1: one (df_worker_cm, A/B)
    int H = cp.Q <= kCm4MaxQ ? 4 : 2;
    if (ep) H = ep[0] == '4' ? (cp.Q <= kCm4MaxQ ? 4 : 2) : ep[0] == '2' ? 2 : 1;
    auto kern = H == 4   ? (P.prof ? df_kernel_cm2<T, UPDATE, true, 4> : df_kernel_cm2<T, UPDATE, false, 4>)
                : H == 2 ? (P.prof ? df_kernel_cm2<T, UPDATE, true, 2> : df_kernel_cm2<T, UPDATE, false, 2>)
                         : (P.prof ? df_kernel_cm<T, UPDATE, true> : df_kernel_cm<T, UPDATE, false>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp.smem);
    if (e != cudaSuccess) return (int)e;
    void *args[] = {(void *)&L, (void *)&v, (void *)&P, (void *)&tmapc, (void *)&tmapcb};
    const int threads = H > 1 ? std::max(kC6ThreadsSplit, 32 * (H * cp.Q + 2)) : kCmMaxThreads;
    if (H == 1 && !make_tmap_cm<T>(&tmapcb, L, n, ld, kPanel * kDfD)) return (int)cudaErrorNotSupported;  // (v6 chain: d tiles)
    e = cudaLaunchCooperativeKernel((void *)kern, dim3(cp.sms), dim3(threads), args, cp.smem, st);
  } else {
    P.R = pl.R; P.r_shift = pl.r_shift; P.Wk = pl.Wk; P.rows_max = pl.rows_max; P.nw = pl.nw;
    auto kern = df_kernel_v2<T, LAYOUT, UPDATE>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return (int)e;
    void *args[] = {(void *)&L, (void *)&v, (void *)&P};
    e = cudaLaunchCooperativeKernel((void *)kern, dim3(pl.sms), dim3(kDfThreads), args, pl.smem, st);
  }
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  if (!UPDATE && !chain_only) {
    if (getenv("CHUB_DD_COEF_V1")) large_dd_coef_kernel<T, LAYOUT><<<1, 1024, 0, st>>>(L, ni, ld, v, ws, status);
    else if (getenv("CHUB_DD_COEF_V2")) dd_coef_par_kernel<T><<<1, 1024, 0, st>>>(ni, v, ws, status);  // (one CTA, A/B)
    else dd_coef_par2_kernel<T><<<(ni + 1023) / 1024, 1024, 0, st>>>(ni, v, ws, status);  // (ws.dg: the pre-pass's copy of the diagonal)
    ++*launches;
    if (rp.ok && !getenv("CHUB_DD_SWEEP_V1") && !getenv("CHUB_DD_SWEEP_V2")) {
      auto sk = dd_sweep_rm2_kernel<T>;
      const size_t ssm = dd_sweep_rm2_smem<T>();
      e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
      if (e != cudaSuccess) return (int)e;
      sk<<<nb, 128, ssm, st>>>(L, ni, ld, ws, status, use3d ? 1 : 0, tmap8, tmap3);
    } else if (rp.ok && !getenv("CHUB_DD_SWEEP_V1")) {
      auto sk = dd_sweep_rm_kernel<T>;
      const size_t ssm = dd_sweep_rm_smem<T>();
      e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
      if (e != cudaSuccess) return (int)e;
      sk<<<nb, 128, ssm, st>>>(L, ni, ld, ws, status, tmap8);
    } else if (cp.ok && !getenv("CHUB_DD_SWEEP_V1")) {
      e = launch_dd_sweep_cm<T>(L, ni, ld, ws, status, tmapc, st);
      if (e != cudaSuccess) return (int)e;
    } else {
      large_dd_sweep_kernel<T, LAYOUT><<<(ni + 127) / 128, 128, 0, st>>>(L, ni, ld, ws, status);
    }
    ++*launches;
  }
  e = cudaGetLastError();
  return (int)e;
}

// ---- blocked driver (row-major): factors beyond the persistent kernel's row capacity.
// Diagonal blocks of at most kBlockedMax rows (a multiple of 32): persistent kernel on block J (input = the
// residual left by the panels to its left, (t, S) carried over), then rect_apply_rm_kernel applies the
// block's columns to every row below it.  A downdate runs the same two kernels read-only (forward
// substitution), then the coefficient and sweep kernels once over the whole factor.
constexpr int64_t kBlockedMax = 16384;

inline LargeWs ws_offset(const LargeWs &ws, int64_t off) {  // the workspace as diagonal block [off, ..) sees it
  LargeWs w = ws;
  w.X += (off / kPanel) * kXBlock;
  if (w.M1) w.M1 += (off / kPanel) * (CHUB_CHAIN_V7 ? 2 : 1) * kXBlock;
  w.w += off; w.p += off; w.c += off; w.sg += off; w.dn += off; w.v0 += off; w.dg += off;
  return w;  // scal and flags are shared
}

// layout-agnostic part of the test; the row-major dispatch (api.cu) uses it for every n beyond one launch
template <typename T>
inline bool blocked_supported(int64_t n, int64_t ld, const T *L, int layout = CHUB_ROW_MAJOR) {
  constexpr int EPC = 16 / sizeof(T);
  if (n <= kBlockedMax || n >= (1LL << 31) || (ld % EPC) != 0 || (reinterpret_cast<uintptr_t>(L) & 15) != 0) return false;
  if (!get_encode_tiled()) return false;
  return layout == CHUB_ROW_MAJOR ? rm_plan<T>(kBlockedMax, ld, L).ok : cm_plan<T>(kBlockedMax, ld, L).ok;
}

template <typename T, int LAYOUT, bool UPDATE>
int launch_blocked(T *L, int64_t n, int64_t ld, T *v, int flags, const LargeWs &ws, int32_t *status,
                   cudaStream_t st, int64_t *launches) {
  *launches = 0;
  const int ni = (int)n;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // block sizes: the fewest blocks that fit one launch each, all about the same size (a multiple of 32).  (A full
  // block followed by a short remainder leaves the rectangle kernel a few hundred row groups walking 16384 columns
  // each: 2.4 TB/s at n = 20000, profiles/r02_ncu_launch_list_n20000.csv.)
  int64_t sizes[64];
  const int64_t nblk64 = (n + kBlockedMax - 1) / kBlockedMax;
  if (nblk64 > 64) return (int)cudaErrorNotSupported;
  const int nblk = (int)nblk64;
  {
    const int64_t each = ((n + nblk - 1) / nblk + kPanel - 1) / kPanel * kPanel;
    int64_t left = n;
    for (int J = 0; J < nblk; ++J) {
      sizes[J] = J + 1 < nblk ? std::min(each, left) : left;
      left -= sizes[J];
    }
  }
  blocked_init_kernel<T, LAYOUT><<<(ni + 255) / 256, 256, 0, st>>>(L, ni, ld, v, ws, flags, status);
  ++*launches;
  // the whole factor as the rectangle kernel (and the downdate's sweep) sees it
  alignas(64) CUtensorMap tmapw;
  if (LAYOUT == CHUB_ROW_MAJOR ? !make_tmap<T>(&tmapw, L, n, ld, 8) : !make_tmap_cm<T>(&tmapw, L, n, ld))
    return (int)cudaErrorNotSupported;
  constexpr bool RM = LAYOUT == CHUB_ROW_MAJOR;
  auto rk = RM ? rect_apply_rm_kernel<T, UPDATE> : rect_apply_cm_kernel<T, UPDATE>;
  const size_t rsm = RM ? rect_apply_rm_smem<T>() : rect_apply_cm_smem<T>();
  const int rwarps = RM ? kRectWarps : kRectCmWarps, rrows = RM ? 8 : kPanel;
  cudaError_t e = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
  if (e != cudaSuccess) return (int)e;
  int64_t off = 0;
  for (int J = 0; J < nblk; ++J) {
    const int64_t nb_ = sizes[J];
    int64_t nl = 0;
    // (check_diag was done for the whole diagonal by the init kernel; status != 0 stops every kernel)
    int rc = launch_dataflow<T, LAYOUT, UPDATE>(L + off * ld + off, nb_, ld, v + off, 0, ws_offset(ws, off),
                                                status, st, &nl, ws.w + off, true);
    *launches += nl;
    if (rc) return rc;
    if (off + nb_ < n) {
      const int rows = (int)(n - off - nb_);
      const int groups = (rows + rrows - 1) / rrows;
      const int grid = std::max(1, RM ? std::min(sms, (groups + rwarps - 1) / rwarps) : std::min(sms, groups));
      rk<<<grid, rwarps * 32, rsm, st>>>(ni, (int)(off + nb_), (int)off, (int)nb_, ws, status, tmapw);
      ++*launches;
    }
    off += nb_;
  }
  if (!UPDATE) {
    const int nb = (ni + 31) / 32;
    if (getenv("CHUB_DD_COEF_V1")) large_dd_coef_kernel<T, LAYOUT><<<1, 1024, 0, st>>>(L, ni, ld, v, ws, status);
    else if (getenv("CHUB_DD_COEF_V2")) dd_coef_par_kernel<T><<<1, 1024, 0, st>>>(ni, v, ws, status);
    else dd_coef_par2_kernel<T><<<(ni + 1023) / 1024, 1024, 0, st>>>(ni, v, ws, status);
    ++*launches;
    if (RM) {
      alignas(64) CUtensorMap tmap3;
      const bool use3d = make_tmap3<T>(&tmap3, L, n, ld, 8, 2 * RmGeom<T>::HALVES);
      if (!use3d) tmap3 = tmapw;
      auto sk = dd_sweep_rm2_kernel<T>;
      const size_t ssm = dd_sweep_rm2_smem<T>();
      e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
      if (e != cudaSuccess) return (int)e;
      sk<<<nb, 128, ssm, st>>>(L, ni, ld, ws, status, use3d ? 1 : 0, tmapw, tmap3);
    } else {
      e = launch_dd_sweep_cm<T>(L, ni, ld, ws, status, tmapw, st);
      if (e != cudaSuccess) return (int)e;
    }
    ++*launches;
  }
  return (int)cudaGetLastError();
}

}  // namespace chub
