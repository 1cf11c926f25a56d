// dataflow_chain_v7.cuh -- chain CTA with ONE mat-vec per block row on the critical path (included by dataflow.cuh).
//
// The barrier-free chain (dataflow_chain_v6.cuh) runs two dependent mat-vecs and two shared-memory hand-overs per
// block row, w_r = g_r - L[r,r-1] p_{r-1} (Ca) and p_r = X_r w_r (Cb): ~1150-1225 cycles per block row, and that
// period IS the pace of the last ~45 % of the block columns of an update and of the whole forward substitution of a
// downdate (chain alone, nobody waiting: 0.299 ms at n = 16384).  Here the pre-pass also forms M1_r = X_r L[r,r-1]
// and M2_r = X_r L[r,r-2] (two 32^3 products per block row, off the kernel), so that
//     p_r = u_r - M2_r p_{r-2} - M1_r p_{r-1},      u_r = X_r g_r,   g_r = wt_r - sum_{j = r-d}^{r-3} L[r,j] p_j
// u_r (band warp -> Cu -> Cp: three hand-overs) depends on nothing newer than p_{r-3}, M2_r p_{r-2} is formed while
// p_{r-1} is still on its way, and the p -> p chain is one poll, 16 broadcast LDS.128 and 32 DFMA.  (With M1 alone
// u_r hangs on p_{r-2} through three hand-overs: 1150 cycles per block row, no better than v6 -- gpurun r04s.)
//   warps 0,1 Cp: p_r (even / odd block rows; M1_r's row in registers before p_{r-1} arrives); publishes p_r
//                (global memory: NaN-sentinel protocol of the workers; shared ring for everybody here)
//   warp 2   Cu : u_r = X_r g_r as soon as g_r is complete, handed over per lane (lane = row: no broadcast needed)
//   warps 3.. B : band tiles at distance 3 .. d, as in v6 (the nearer two are not loaded: M1 / M2 stand for them)
//   then     P  : polls wt_r from global memory into the shared ring
//   last     T  : one TMA box per block column (d-1 band tiles) + X_{j+1} + M1_{j+1} into the slot ring
#pragma once

namespace chub {

constexpr int kC7NT = kDfD - 2;  // band tiles per slot (distances 3 .. d)
constexpr int kC7Threads = 32 * (5 + kC7NT);  // Cp x 2, Cu, band warps, P, T

template <typename T>
struct C7Geom {
  static constexpr int TILES = kC7NT * RmGeom<T>::CTILE;
  static constexpr int SLOT = ((TILES + 3 * kXBlock * 8 + 1023) / 1024) * 1024;  // + X_{j+1} + M1_{j+1} + M2_{j+1}
  static constexpr int BASE = 10240;  // control block + rings
};

// Row `lane` of band tile q (distance q + 3) of a ring slot, into registers.
template <typename T, int LAYOUT>
__device__ __forceinline__ void c7_tile_row(const unsigned char *slot, int q, int lane, double (&t)[kPanel]) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, EPC = RmGeom<T>::EPC;
  if (LAYOUT == CHUB_COL_MAJOR) {
    const T *tc = reinterpret_cast<const T *>(slot) + q * kPanel + lane;
#pragma unroll
    for (int k = 0; k < kPanel; ++k) t[k] = (double)tc[k * (kC7NT * kPanel)];
  } else {
#pragma unroll
    for (int h = 0; h < HALVES; ++h) {
      const unsigned char *rowp = slot + h * (kC7NT * 4096) + q * 4096 + lane * 128;
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

template <typename T, int LAYOUT, bool PROF>
__device__ void df_chain_v7(const DfParams &P, const CUtensorMap *tmapB, unsigned char *smem) {
  constexpr int HALVES = RmGeom<T>::HALVES;
  constexpr int TILES = C7Geom<T>::TILES, SLOT = C7Geom<T>::SLOT;
  constexpr int NB = kC7NT;  // band warps: distances 3 .. d, block row r belongs to band warp (r - 3) mod NB
  static_assert(NB >= 1 && 5 + NB <= 16, "warp roles");
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int T_ = P.T;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [kDfSlots]
  volatile int *ctl = reinterpret_cast<volatile int *>(smem + 64);
  volatile int *pdone = ctl;        // p_r is in pbuf for r < *pdone
  volatile int *udone = ctl + 1;    // u_r has been formed for r < *udone (wt_r consumed)
  volatile int *wtdone = ctl + 2;   // wt_r is in wtb for r < *wtdone
  volatile int *carel = ctl + 24;   // [2] block columns whose M1 Cp (even / odd rows) holds in registers
  volatile int *cbrel = ctl + 4;    // block columns whose X Cu holds in registers
  volatile int *bprog = ctl + 8;    // [4] block columns band warp k has finished reading
  volatile int *gready = ctl + 16;  // [kC6Ring] r + 1 once the band part of block row r is in gbuf[r % ring]
  double *pbuf = reinterpret_cast<double *>(smem + 512);   // [kC6Ring][32]
  double *gbuf = pbuf + kC6Ring * kPanel;                   // [kC6Ring][32]
  double *wtb = gbuf + kC6Ring * kPanel;                    // [kC6Ring][32]
  double *ubuf = wtb + kC6Ring * kPanel;                    // [kC6Ring][32]  u_r, lane = row (sentinel = not there)
  double *gvec = ubuf + kC6Ring * kPanel;                   // [2][32]        g_r for Cu's broadcast reads
  unsigned char *slots = smem + C7Geom<T>::BASE;            // [kDfSlots][SLOT]
  static_assert(512 + (4 * kC6Ring + 2) * kPanel * 8 <= C7Geom<T>::BASE, "control block");

  if (tid == 0) {
    for (int i = 0; i < kDfSlots; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  if (tid < 64) ctl[tid] = 0;
  {
    const double sentinel = __longlong_as_double((long long)kSentinelBits);
    for (int e = tid; e < kC6Ring * kPanel; e += kC7Threads) { pbuf[e] = sentinel; ubuf[e] = sentinel; }
  }
  named_sync(1, kC7Threads);

  const int late0 = (PROF && P.prof && P.prof_late) ? T_ * 5 / 8 : -1;

  if (warp < 2) {
    // ------------------------------------------------------------ Cp: p_r = u_r - M1_r p_{r-1}
    long long t_chain0 = DF_PROF_T();
    for (int r = warp; r < T_; r += 2) {
      if (late0 >= 0 && (r == late0 || r == late0 + 1)) {
        for (int i = 0; i < 8; ++i) pacc[i] = 0;
        t_chain0 = clock64();
      }
      const int j = r - 1;
      double t[kPanel];
      double acc2 = 0.0;  // M2_r p_{r-2}
      if (r > 0) {
        long long tp = DF_PROF_T();
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        if (lane == 0) DF_PROF_ADD(1, tp);
        const double *M = reinterpret_cast<const double *>(slots + (size_t)(j % kDfSlots) * SLOT + TILES + kXBlock * 8) + lane * kXStride;
        if (r > 1) {  // the distance-2 term, while p_{r-1} is still on its way
#pragma unroll
          for (int k = 0; k < kPanel; k += 2) {
            const double2 a = *reinterpret_cast<const double2 *>(M + kXBlock + k);
            t[k] = a.x; t[k + 1] = a.y;
          }
          c6_poll_vec(pbuf + ((r - 2) % kC6Ring) * kPanel, lane, P.status, false);  // p_{r-2}: normally there
          acc2 = c6_dot(t, pbuf + ((r - 2) % kC6Ring) * kPanel);
        }
#pragma unroll
        for (int k = 0; k < kPanel; k += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(M + k);
          t[k] = a.x; t[k + 1] = a.y;
        }
        __syncwarp();
        if (lane == 0) c6_store_flag_after(carel + warp, j + 1, t[kPanel - 1], t[kPanel / 2 - 1]);
      }
      // ring entry of p_{r-5} -> sentinel (it becomes p_{r+3}): every reader is past it -- this block row's TMA
      // box was only issued once the band warps had finished block column r-5
      c6_store_vec(pbuf + ((r + 3) % kC6Ring) * kPanel, lane, __longlong_as_double((long long)kSentinelBits));
      // u_r (mine: lane = row); normally there long before p_{r-1}
      long long tw = DF_PROF_T();
      volatile unsigned long long *up = reinterpret_cast<volatile unsigned long long *>(ubuf + (r % kC6Ring) * kPanel) + lane;
      unsigned long long ub = *up;
      if (ub == kSentinelBits) {
        const long long t0 = clock64();
        while ((ub = *up) == kSentinelBits) {
          if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); ub = 0; break; }
        }
      }
      *up = kSentinelBits;  // taken
      if (lane == 0) DF_PROF_ADD(3, tw);
      double pr = __longlong_as_double((long long)ub) - acc2;
      long long tp = DF_PROF_T();
      if (r > 0) {
        c6_poll_vec(pbuf + (j % kC6Ring) * kPanel, lane, P.status, false);  // p_{r-1}
        if (lane == 0) DF_PROF_ADD(2, tp);
        tp = DF_PROF_T();
        pr -= c6_dot(t, pbuf + (j % kC6Ring) * kPanel);
      }
      st_relaxed_f64(P.ws.p + (int64_t)r * kPanel + lane, pr);  // publish: the data is its own flag
      c6_store_vec(pbuf + (r % kC6Ring) * kPanel, lane, df_sanitize(pr));
      if (lane == 0) {
        *pdone = r + 1;
        DF_PROF_ADD(5, tp);
        DF_TRACE(2, r);
      }
    }
    if (lane == 0 && warp == 0) DF_PROF_ADD(0, t_chain0);
  } else if (warp == 2) {
    // ------------------------------------------------------------ Cu: u_r = X_r g_r
    for (int r = 0; r < T_; ++r) {
      double x[kPanel];
      if (r == 0) {
        const double *X0 = P.ws.X + lane * kXStride;
#pragma unroll
        for (int k = 0; k < kPanel; ++k) x[k] = X0[k];
      } else {
        const int j = r - 1;
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        const double *X = reinterpret_cast<const double *>(slots + (size_t)(j % kDfSlots) * SLOT + TILES) + lane * kXStride;
#pragma unroll
        for (int k = 0; k < kPanel; k += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(X + k);
          x[k] = a.x; x[k + 1] = a.y;
        }
        __syncwarp();
        if (lane == 0) c6_store_flag_after(cbrel, j + 1, x[kPanel - 1], x[kPanel / 2 - 1]);
      }
      c6_spin(wtdone, r + 1, P.status);
      double g = wtb[(r % kC6Ring) * kPanel + lane];
      if (r >= 3) {
        c6_spin(gready + (r % kC6Ring), r + 1, P.status);
        g += gbuf[(r % kC6Ring) * kPanel + lane];
      }
      double *gv = gvec + (r & 1) * kPanel;
      gv[lane] = g;
      __syncwarp();
      const double u = c6_dot(x, gv);
      // (the slot's previous tenant u_{r-8} was taken long ago: p_{r-2} exists)
      c6_store_vec(ubuf + (r % kC6Ring) * kPanel, lane, df_sanitize(u));
      if (lane == 0) *udone = r + 1;
    }
  } else if (warp < 3 + NB) {
    // ------------------------------------------------------------ B: band tiles at distance 3 .. d
    const int k4 = warp - 3;
    double acc = 0.0;
    for (int j = 0; j < T_; ++j) {
      const int r = j + 3 + (((k4 - j) % NB + NB) % NB);  // my block row among j+3 .. j+d:  r = k4 + 3 (mod NB)
      if (r < T_) {
        if (j == 0 || r - j == kDfD) acc = 0.0;  // first tile of this block row
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        double t[kPanel];
        c7_tile_row<T, LAYOUT>(slots + (size_t)(j % kDfSlots) * SLOT, r - j - 3, lane, t);
        c6_poll_vec(pbuf + (j % kC6Ring) * kPanel, lane, P.status, CHUB_C6_LAZY != 0);
        acc -= c6_dot(t, pbuf + (j % kC6Ring) * kPanel);
        if (r - j == 3) {  // last band tile of block row r (gbuf[r % 8] was read for u_{r-8}: long past)
          gbuf[(r % kC6Ring) * kPanel + lane] = acc;
          __syncwarp();
          if (lane == 0) st_release_cta_smem(gready + (r % kC6Ring), r + 1);
        }
      }
      __syncwarp();
      if (lane == 0) st_release_cta_smem(bprog + k4, j + 1);  // my reads of ring slot j % kDfSlots are done
    }
  } else if (warp == 3 + NB) {
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
      if (r >= kC6Ring) c6_spin_lazy(udone, r - kC6Ring + 1, P.status);  // ring slot: wt_{r-8} has been consumed
      wtb[(r % kC6Ring) * kPanel + lane] = wt;
      __syncwarp();
      if (lane == 0) st_release_cta_smem(wtdone, r + 1);
    }
  } else {
    // ------------------------------------------------------------ T: the TMA boxes
    for (int jj = 0; jj + 1 < T_; ++jj) {
      if (elect_one()) {
        if (kDfPf > 0 && jj + kDfPf + 3 < T_) {  // the band box of a later block column -> L2
          if (LAYOUT == CHUB_COL_MAJOR) tma_prefetch_2d(tmapB, (jj + kDfPf + 3) * kPanel, (jj + kDfPf) * kPanel);
          else tma_prefetch_3d(tmapB, 0, (jj + kDfPf + 3) * kPanel, (jj + kDfPf) * HALVES);
        }
        if (jj >= kDfSlots) {  // the slot held block column jj - slots: taken by Cp (M1), Cu (X), the band warps
          const int done = jj - kDfSlots + 1;
          c6_spin_lazy(carel + (done & 1), done, P.status);  // (block column done-1 = block row done; the other
                                                             //  parity's previous column was checked one jj ago)
          c6_spin_lazy(cbrel, done, P.status);
          for (int k = 0; k < NB; ++k) c6_spin_lazy(bprog + k, done, P.status);
        }
        const int slot = jj % kDfSlots;
        void *bar = &full[slot];
        unsigned char *sb = slots + (size_t)slot * SLOT;
        const bool has_band = jj + 3 < T_;  // (the last block columns have nothing below their distance-2 tile)
        mbar_arrive_expect_tx(bar, (has_band ? (unsigned)TILES : 0u) + 3u * kXBlock * 8u);
        // ONE box: block rows jj+3 .. jj+d of block column jj (rows past the end of the matrix are zero-filled)
        if (has_band) {
          if (LAYOUT == CHUB_COL_MAJOR) tma_load_2d(sb, tmapB, (jj + 3) * kPanel, jj * kPanel, bar);
          else tma_load_3d(sb, tmapB, 0, (jj + 3) * kPanel, jj * HALVES, bar);
        }
        bulk_g2s(sb + TILES, P.ws.X + (int64_t)(jj + 1) * kXBlock, kXBlock * 8u, bar);
        bulk_g2s(sb + TILES + kXBlock * 8, P.ws.M1 + (int64_t)(jj + 1) * 2 * kXBlock, 2u * kXBlock * 8u, bar);  // [M1 | M2]
      }
      __syncwarp();
    }
    return;
  }
  if (PROF && P.prof && lane == 0 && (warp == 0 || warp == 2 || warp == 3 + NB)) {  // Cp (even rows), Cu, P
    for (int i = 0; i < 8; ++i)
      if (pacc[i]) P.prof[(size_t)blockIdx.x * 16 + i + (warp == 2 ? 8 : 0)] = pacc[i];
  }
}

template <typename T>
inline size_t df_chain_v7_smem() {
  return (size_t)C7Geom<T>::BASE + (size_t)kDfSlots * C7Geom<T>::SLOT;
}

}  // namespace chub
