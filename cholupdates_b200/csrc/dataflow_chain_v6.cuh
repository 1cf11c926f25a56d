// dataflow_chain_v6.cuh -- chain CTA of the persistent kernels without CTA-wide barriers (included by dataflow.cuh).
//
// Why: the barrier chain (df_chain_rm) takes ~1430 cycles per block row even when nobody makes it wait
// (profiles/r03a): five band mat-vecs, a barrier, the X mat-vec on ONE warp, another barrier -- and in the last
// 45 % of the block columns, where the remaining triangle is too small to keep HBM busy, that period IS the
// kernel's pace.  Here the eight warps have fixed roles and hand over through shared-memory counters
// (st.release.cta / ld.acquire.cta, every spin with a watchdog):
//   warps 0,1 Ca: w_r = g_r - L[r, r-1] p_{r-1} (even / odd block rows).  The tile L[r, r-1] is already in
//                REGISTERS when p_{r-1} arrives (loaded as soon as the TMA box of block column r-1 lands) and
//                g_r = wt_r + band part is in hand: with two warps the preamble of block row r+1 (barrier wait,
//                16 LDS.128, flags) runs while the other warp is on the critical path of block row r.
//   warp 2   Cb : p_r = X_r w_r, X_r in registers ahead of time; publishes p_r (global memory: NaN-sentinel
//                protocol of the workers; shared ring for the other warps).
//   warps 3.. B : the band tiles at distance 2 .. d: block row r belongs to band warp (r - 2) mod (d-1), which adds
//                L[r, j] p_j for j = r-d .. r-2 as the p_j arrive (accumulator in registers, lane = row) and
//                hands the sum over one block row before it is needed.
//   last     PT : polls wt_r (handed over by the workers through global memory) into a shared ring, and issues one
//                TMA box per block column (all d band tiles) + X as soon as the ring slot's previous block column
//                has been taken by everybody; L2 prefetch further ahead.
// The two hand-overs on the critical path (w_r: Ca -> Cb, p_r: Cb -> Ca) carry no flag and no fence: the data
// is its own flag (NaN sentinel in the shared-memory ring, lane l polls element l, __syncwarp, then everybody
// reads all 32).  The critical path per block row is: poll -> 16 broadcast LDS -> 32 DFMA (4 chains), twice.
#pragma once

namespace chub {

constexpr int kC6Ring = 8;  // p / g / wt rings (block rows)
constexpr int kC6ThreadsMerged = 32 * (4 + kDfD - 1);  // Ca x 2, Cb, d-1 band warps, PT
constexpr int kC6ThreadsSplit = 32 * (5 + kDfD - 1);   // ... with P and T on a warp each

template <typename T>
struct C6Geom {
  static constexpr int TILES = kDfD * RmGeom<T>::CTILE;                       // the band box of a block column
  static constexpr int SLOT = ((TILES + kXBlock * 8 + 1023) / 1024) * 1024;   // + X_{j+1}
  static constexpr int BASE = 8192;                                           // control block + rings
};

__device__ __forceinline__ void c6_spin(const volatile int *flag, int need, int32_t *status) {
  if (ld_acquire_cta_smem(flag) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire_cta_smem(flag) < need) {
    if (clock64() - t0 > kWatchdogCycles) { atomicMax(status, CHUB_STATUS_INTERNAL); return; }
  }
}

// Data-as-flag hand-over inside the CTA: lane l spins until element l of a 32-vector in shared memory is no
// longer the sentinel; after the __syncwarp every lane may read all 32 elements.
__device__ __forceinline__ void c6_poll_vec(const double *vec, int lane, int32_t *status, bool lazy) {
  const volatile unsigned long long *q = reinterpret_cast<const volatile unsigned long long *>(vec) + lane;
  if (*q == kSentinelBits) {
    const long long t0 = clock64();
    while (*q == kSentinelBits) {
      if (lazy) __nanosleep(40);
      if (clock64() - t0 > kWatchdogCycles) { atomicMax(status, CHUB_STATUS_INTERNAL); break; }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void c6_store_vec(double *vec, int lane, double x) {
  *(reinterpret_cast<volatile double *>(vec) + lane) = x;
}

// Plain flag store that issues only once the registers a, b have been loaded (the last shared-memory loads of a
// register prefetch): "my reads of this ring slot are complete" without a fence -- a st.release here would
// wait for the warp's global-memory store of p (hundreds of cycles, measured on Cb's path: profiles/r03f).
__device__ __forceinline__ void c6_store_flag_after(volatile int *flag, int x, double a, double b) {
  asm volatile("{\n\t.reg .f64 t;\n\tadd.f64 t, %2, %3;\n\tst.volatile.shared.s32 [%0], %1;\n\t}\n"
               ::"r"(smem_u32(const_cast<int *>(flag))), "r"(x), "d"(a), "d"(b) : "memory");
}

// The same for the warps that are NOT on the critical path (band warps, T, P's ring check): back off between
// polls so that their shared-memory loads do not compete with Ca / Cb for the LSU and the issue slots.
#ifndef CHUB_C6_LAZY
#define CHUB_C6_LAZY 1
#endif
__device__ __forceinline__ void c6_spin_lazy(const volatile int *flag, int need, int32_t *status) {
  if (ld_acquire_cta_smem(flag) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire_cta_smem(flag) < need) {
    if (CHUB_C6_LAZY) __nanosleep(40);
    if (clock64() - t0 > kWatchdogCycles) { atomicMax(status, CHUB_STATUS_INTERNAL); return; }
  }
}

// Row `lane` of band tile q of a ring slot, into registers.  Slot layouts: row-major source
// [half][tile q][32 rows][128 B, SWIZZLE_128B]; column-major source [32 columns][32 d rows].
template <typename T, int LAYOUT>
__device__ __forceinline__ void c6_tile_row(const unsigned char *slot, int q, int lane, double (&t)[kPanel]) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, EPC = RmGeom<T>::EPC;
  if (LAYOUT == CHUB_COL_MAJOR) {
    const T *tc = reinterpret_cast<const T *>(slot) + q * kPanel + lane;
#pragma unroll
    for (int k = 0; k < kPanel; ++k) t[k] = (double)tc[k * (kDfD * kPanel)];
  } else {
#pragma unroll
    for (int h = 0; h < HALVES; ++h) {
      const unsigned char *rowp = slot + h * (kDfD * 4096) + q * 4096 + lane * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int4 raw = *reinterpret_cast<const int4 *>(rowp + ((c ^ (lane & 7)) << 4));
        const T *x = reinterpret_cast<const T *>(&raw);
#pragma unroll
        for (int u = 0; u < EPC; ++u) t[h * BOXC + c * EPC + u] = (double)x[u];
      }
    }
  }
}

// sum_k t[k] vec[k], vec in shared memory (all lanes read the same address: broadcast), four chains
__device__ __forceinline__ double c6_dot(const double (&t)[kPanel], const double *vec) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int k = 0; k < kPanel; k += 4) {
    const double2 u = *reinterpret_cast<const double2 *>(vec + k);
    const double2 z = *reinterpret_cast<const double2 *>(vec + k + 2);
    a0 = fma(t[k], u.x, a0);
    a1 = fma(t[k + 1], u.y, a1);
    a2 = fma(t[k + 2], z.x, a2);
    a3 = fma(t[k + 3], z.y, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// MERGE_PT: one warp both polls wt and issues the TMA boxes (8 warps for d = 5: the column-major workers'
// register budget allows 256 threads); split, the row-major update is 10 % faster (0.452 vs 0.50 ms at n = 16384:
// the poller never sits behind a ring-slot wait, the TMA issue never behind a poll).
template <typename T, int LAYOUT, bool PROF, bool MERGE_PT>
__device__ void df_chain_v6(const DfParams &P, const CUtensorMap *tmapB, unsigned char *smem) {
  constexpr int kC6Threads = MERGE_PT ? kC6ThreadsMerged : kC6ThreadsSplit;
  constexpr int HALVES = RmGeom<T>::HALVES;
  constexpr int TILES = C6Geom<T>::TILES, SLOT = C6Geom<T>::SLOT;
  constexpr int NB = kDfD - 1;  // band warps: distances 2 .. d, block row r belongs to band warp (r - 2) mod NB
  static_assert(NB >= 1 && 4 + NB <= 16, "warp roles");
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int T_ = P.T;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [kDfSlots]
  volatile int *ctl = reinterpret_cast<volatile int *>(smem + 64);
  volatile int *pdone = ctl;        // p_r is in pbuf for r < *pdone
  volatile int *wdone = ctl + 1;    // w_r is in wbuf for r < *wdone
  volatile int *wtdone = ctl + 2;   // wt_r is in wtb for r < *wtdone
  volatile int *carel = ctl + 24;   // [2] block columns whose distance-1 tile Ca (even / odd rows) holds in registers
  volatile int *cbrel = ctl + 4;    // block columns whose X Cb holds in registers
  volatile int *bprog = ctl + 8;    // [4] block columns band warp k has finished reading
  volatile int *gready = ctl + 16;  // [kC6Ring] r + 1 once the band part of block row r is in gbuf[r % ring]
  double *pbuf = reinterpret_cast<double *>(smem + 512);   // [kC6Ring][32]
  double *gbuf = pbuf + kC6Ring * kPanel;                   // [kC6Ring][32]
  double *wtb = gbuf + kC6Ring * kPanel;                    // [kC6Ring][32]
  double *wbuf = wtb + kC6Ring * kPanel;                    // [2][32]
  unsigned char *slots = smem + C6Geom<T>::BASE;            // [kDfSlots][SLOT]
  static_assert(512 + (3 * kC6Ring + 2) * kPanel * 8 <= C6Geom<T>::BASE, "control block");

  if (tid == 0) {
    for (int i = 0; i < kDfSlots; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  if (tid < 64) ctl[tid] = 0;
  {
    const double sentinel = __longlong_as_double((long long)kSentinelBits);
    for (int e = tid; e < kC6Ring * kPanel; e += kC6Threads) pbuf[e] = sentinel;
    if (tid < 2 * kPanel) wbuf[tid] = sentinel;
  }
  named_sync(1, kC6Threads);

  const int late0 = (PROF && P.prof && P.prof_late) ? T_ * 5 / 8 : -1;

  if (warp < 2) {
    // ------------------------------------------------------------ Ca: w_r = wt_r + band_r - L[r, r-1] p_{r-1}
    for (int r = warp; r < T_; r += 2) {
      if (late0 >= 0 && (r == late0 || r == late0 + 1)) for (int i = 0; i < 8; ++i) pacc[i] = 0;
      const int j = r - 1;
      double t[kPanel];
      if (r > 0) {
        long long tp = DF_PROF_T();
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        if (lane == 0) DF_PROF_ADD(1, tp);
        c6_tile_row<T, LAYOUT>(slots + (size_t)(j % kDfSlots) * SLOT, 0, lane, t);
        __syncwarp();
        if (lane == 0) c6_store_flag_after(carel + warp, j + 1, t[kPanel - 1], t[kPanel / 2 - 1]);
      }
      // g_r: both parts are there long before p_{r-1}
      long long tw = DF_PROF_T();
      c6_spin(wtdone, r + 1, P.status);
      if (lane == 0) DF_PROF_ADD(3, tw);
      double g = wtb[(r % kC6Ring) * kPanel + lane];
      if (r >= 2) {
        c6_spin(gready + (r % kC6Ring), r + 1, P.status);
        g += gbuf[(r % kC6Ring) * kPanel + lane];
      }
      if (r > 0) {
        long long tp = DF_PROF_T();
        c6_poll_vec(pbuf + (j % kC6Ring) * kPanel, lane, P.status, false);  // p_{r-1}
        if (lane == 0) DF_PROF_ADD(2, tp);
        tp = DF_PROF_T();
        g -= c6_dot(t, pbuf + (j % kC6Ring) * kPanel);
        c6_store_vec(wbuf + (r & 1) * kPanel, lane, df_sanitize(g));  // (Cb reset this slot after p_{r-2})
        if (lane == 0) { *wdone = r + 1; DF_PROF_ADD(4, tp); }
      } else {
        c6_store_vec(wbuf, lane, df_sanitize(g));
        if (lane == 0) *wdone = 1;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ Cb: p_r = X_r w_r
    long long t_chain0 = DF_PROF_T();
    for (int r = 0; r < T_; ++r) {
      if (r == late0) {
        for (int i = 0; i < 8; ++i) pacc[i] = 0;
        t_chain0 = clock64();
      }
      double x[kPanel];
      if (r == 0) {
        const double *X0 = P.ws.X + lane * kXStride;
#pragma unroll
        for (int k = 0; k < kPanel; ++k) x[k] = X0[k];
      } else {
        const int j = r - 1;
        long long tp = DF_PROF_T();
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        if (lane == 0) DF_PROF_ADD(1, tp);
        const double *X = reinterpret_cast<const double *>(slots + (size_t)(j % kDfSlots) * SLOT + TILES) + lane * kXStride;
#pragma unroll
        for (int k = 0; k < kPanel; k += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(X + k);
          x[k] = a.x; x[k + 1] = a.y;
        }
        __syncwarp();
        if (lane == 0) c6_store_flag_after(cbrel, j + 1, x[kPanel - 1], x[kPanel / 2 - 1]);
      }
      // ring entry of p_{r-5} -> sentinel (it becomes p_{r+3}): every reader is past it -- this block row's TMA
      // box was only issued once the band warps had finished block column r-5
      c6_store_vec(pbuf + ((r + 3) % kC6Ring) * kPanel, lane, __longlong_as_double((long long)kSentinelBits));
      long long tp = DF_PROF_T();
      c6_poll_vec(wbuf + (r & 1) * kPanel, lane, P.status, false);
      if (lane == 0) DF_PROF_ADD(3, tp);
      tp = DF_PROF_T();
      const double pr = c6_dot(x, wbuf + (r & 1) * kPanel);
      st_relaxed_f64(P.ws.p + (int64_t)r * kPanel + lane, pr);  // publish: the data is its own flag
      c6_store_vec(pbuf + (r % kC6Ring) * kPanel, lane, df_sanitize(pr));
      __syncwarp();  // every lane has read w_r
      c6_store_vec(wbuf + (r & 1) * kPanel, lane, __longlong_as_double((long long)kSentinelBits));
      if (lane == 0) {
        *pdone = r + 1;
        DF_PROF_ADD(5, tp);
        DF_TRACE(2, r);
      }
    }
    if (lane == 0) DF_PROF_ADD(0, t_chain0);
  } else if (warp < 3 + NB) {
    // ------------------------------------------------------------ B: band tiles at distance 2 .. d
    const int k4 = warp - 3;
    double acc = 0.0;
    for (int j = 0; j < T_; ++j) {
      const int r = j + 2 + (((k4 - j) % NB + NB) % NB);  // my block row among j+2 .. j+d:  r = k4 + 2 (mod NB)
      if (r < T_) {
        if (j == 0 || r - j == kDfD) acc = 0.0;  // first tile of this block row
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        double t[kPanel];
        c6_tile_row<T, LAYOUT>(slots + (size_t)(j % kDfSlots) * SLOT, r - j - 1, lane, t);
        c6_poll_vec(pbuf + (j % kC6Ring) * kPanel, lane, P.status, CHUB_C6_LAZY != 0);
        acc -= c6_dot(t, pbuf + (j % kC6Ring) * kPanel);
        if (r - j == 2) {  // last band tile of block row r (gbuf[r % 8] was read for w_{r-8}: long past)
          gbuf[(r % kC6Ring) * kPanel + lane] = acc;
          __syncwarp();
          if (lane == 0) st_release_cta_smem(gready + (r % kC6Ring), r + 1);
        }
      }
      __syncwarp();
      if (lane == 0) st_release_cta_smem(bprog + k4, j + 1);  // my reads of ring slot j % kDfSlots are done
    }
  } else if (!MERGE_PT && warp == 3 + NB) {
    // ------------------------------------------------------------ P: wt_r from the workers (global memory)
    unsigned long long b0 = ld_relaxed_u64(P.ws.w + lane);
    for (int r = 0; r < T_; ++r) {
      if (r == late0) pacc[4] = 0;
      long long tq = DF_PROF_T();
#ifdef CHUB_EXP_CHAIN_FREE  // timing experiment (results are garbage): the chain never waits for the workers
      const double wt = b0 != kSentinelBits ? __longlong_as_double((long long)b0) : 0.0;
#else
      const double wt = b0 != kSentinelBits ? __longlong_as_double((long long)b0)
                                            : df_wait_value(P.ws.w + (int64_t)r * kPanel + lane, P.status);
#endif
      b0 = r + 1 < T_ ? ld_relaxed_u64(P.ws.w + (int64_t)(r + 1) * kPanel + lane) : kSentinelBits;
      if (lane == 0) { DF_PROF_ADD(4, tq); DF_TRACE(1, r); }
      if (r >= kC6Ring) c6_spin_lazy(wdone, r - kC6Ring + 1, P.status);  // ring slot: w_{r-8} has been formed
      wtb[(r % kC6Ring) * kPanel + lane] = wt;
      __syncwarp();
      if (lane == 0) st_release_cta_smem(wtdone, r + 1);
    }
  } else {
    // ------------------------------------------------------------ T (split) / PT (merged)
    // PT: wt_r from the workers (global memory) into
    // the shared ring, and the TMA box of block column r + slots-1 (one warp: the CTA stays at 8 warps for d = 5,
    // which the column-major workers' register budget needs).  Order: P(r) then T(r + slots - 1) -- T waits for
    // block column r-1 to have been consumed, which needs wt_{r-1}: delivered one iteration earlier.
    auto issue_column = [&](int jj) {  // all lanes call; one elected lane works
      if (jj + 1 >= T_) return;
      if (elect_one()) {
        if (kDfPf > 0 && jj + kDfPf + 1 < T_) {  // the band box of a later block column -> L2
          if (LAYOUT == CHUB_COL_MAJOR) tma_prefetch_2d(tmapB, (jj + kDfPf + 1) * kPanel, (jj + kDfPf) * kPanel);
          else tma_prefetch_3d(tmapB, 0, (jj + kDfPf + 1) * kPanel, (jj + kDfPf) * HALVES);
        }
        if (jj >= kDfSlots) {  // the slot held block column jj - slots: taken by Ca (tile), Cb (X), the band warps
          const int done = jj - kDfSlots + 1;
          c6_spin_lazy(carel + (done & 1), done, P.status);  // (block column done-1 = block row done; the other
                                                             //  parity's previous column was checked one jj ago)
          c6_spin_lazy(cbrel, done, P.status);
          for (int k = 0; k < NB; ++k) c6_spin_lazy(bprog + k, done, P.status);
        }
        const int slot = jj % kDfSlots;
        void *bar = &full[slot];
        unsigned char *sb = slots + (size_t)slot * SLOT;
        mbar_arrive_expect_tx(bar, (unsigned)TILES + kXBlock * 8u);
        // ONE box: block rows jj+1 .. jj+d of block column jj (rows past the end of the matrix are zero-filled)
        if (LAYOUT == CHUB_COL_MAJOR) tma_load_2d(sb, tmapB, (jj + 1) * kPanel, jj * kPanel, bar);
        else tma_load_3d(sb, tmapB, 0, (jj + 1) * kPanel, jj * HALVES, bar);
        bulk_g2s(sb + TILES, P.ws.X + (int64_t)(jj + 1) * kXBlock, kXBlock * 8u, bar);
      }
      __syncwarp();
    };
    if (!MERGE_PT) {
      for (int jj = 0; jj + 1 < T_; ++jj) issue_column(jj);
      return;
    }
    for (int jj = 0; jj < kDfSlots - 1; ++jj) issue_column(jj);
    unsigned long long b0 = ld_relaxed_u64(P.ws.w + lane);
    for (int r = 0; r < T_; ++r) {
      if (r == late0) pacc[4] = 0;
      long long tq = DF_PROF_T();
#ifdef CHUB_EXP_CHAIN_FREE  // timing experiment (results are garbage): the chain never waits for the workers
      const double wt = b0 != kSentinelBits ? __longlong_as_double((long long)b0) : 0.0;
#else
      const double wt = b0 != kSentinelBits ? __longlong_as_double((long long)b0)
                                            : df_wait_value(P.ws.w + (int64_t)r * kPanel + lane, P.status);
#endif
      b0 = r + 1 < T_ ? ld_relaxed_u64(P.ws.w + (int64_t)(r + 1) * kPanel + lane) : kSentinelBits;
      if (lane == 0) { DF_PROF_ADD(4, tq); DF_TRACE(1, r); }
      if (r >= kC6Ring) c6_spin_lazy(wdone, r - kC6Ring + 1, P.status);  // ring slot: w_{r-8} has been formed
      wtb[(r % kC6Ring) * kPanel + lane] = wt;
      __syncwarp();
      if (lane == 0) st_release_cta_smem(wtdone, r + 1);
      issue_column(r + kDfSlots - 1);
    }
  }
  if (PROF && P.prof && lane == 0 && (warp == 0 || warp == 2 || warp == 3 + NB)) {  // Ca (even rows), Cb, PT
    for (int i = 0; i < 8; ++i)
      if (pacc[i]) P.prof[(size_t)blockIdx.x * 16 + i + (warp == 0 ? 8 : 0)] = pacc[i];
  }
}

template <typename T>
inline size_t df_chain_v6_smem() {
  return (size_t)C6Geom<T>::BASE + (size_t)kDfSlots * C6Geom<T>::SLOT;
}

}  // namespace chub
