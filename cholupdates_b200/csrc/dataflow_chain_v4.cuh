// dataflow_chain_v4.cuh -- EXPERIMENT, not compiled into the library (kept as a measured dead end).
//
// A chain CTA without CTA-wide barriers: fixed warp roles, hand-offs through shared-memory counters, and
// the block step p_{r-1} -> p_r reduced to ONE 32x32 mat-vec by folding the nearest band tile into
// M1_r = X_r L[r, r-1] (formed by the pre-pass; large.cuh still computes it on request).  The chain's own
// time per block row drops from ~1750 to a few hundred cycles, but the update does not get faster: the
// period of the persistent kernel is set by the chain <-> worker hand-over loop (two global-memory hops
// of ~2 us each under load, divided by the look-ahead d + 1), not by the chain's arithmetic.  Measured
// (B200, n = 16384 fp64, same box): barrier chain 0.465 ms, this chain 0.488-0.534 ms depending on where
// the coefficients are formed (profiles/r02_chain_v4_experiments.txt).  To use it, include this header
// after dataflow.cuh's primitives, call df_chain_v4 from the kernels and launch the pre-pass with
// with_m1 = true.
#pragma once

namespace chub {

// ---- chain CTA (threads 0..255 of CTA 0): eight warps with fixed roles, no CTA-wide barrier in the loop.
//
// With M1_r = X_r L[r, r-1] (formed by the pre-pass, off the critical path) the forward substitution reads
//     p_r = ua_r + xw_r - M1_r p_{r-1},   ua_r = -X_r sum_{q=2..d} L[r, r-q] p_{r-q},   xw_r = X_r wt_r
// so the step p_{r-1} -> p_r is ONE 32x32 mat-vec (warp C).  ua_r is complete one step earlier: block row r
// belongs to band warp (r - 2) mod 4, which adds the band tiles (r, j), j = r-d .. r-2, as the p_j arrive
// (one tile per step, accumulator in registers, lane = row) and multiplies by X_r after the last one.
//   warp 0      C: p_r = u_r - M1_r p_{r-1}; publishes p_r (shared ring + global)
//   warps 1..4  B: band tiles + u_r
//   warp 5      P: polls wt_r (handed over by the workers) and forms xw_r = X_r wt_r, so that wt_r is needed
//                  only one mat-vec before p_r is due:  p_r = ua_r + xw_r - M1_r p_{r-1}
//   warp 6      T: issues the TMA loads of a block column as soon as its ring slot has been consumed, and
//                  prefetches the band tiles of a later block column into L2
// Hand-offs are counters in shared memory (st.release.cta / ld.acquire.cta); every spin has a watchdog.
constexpr int kChTiles = kDfD - 1;  // band tiles per block column in the ring (distance 1 is folded into M1)
constexpr int kChRing = 8;          // p / u / wt rings (block rows)

template <typename T>
struct ChGeom {
  static constexpr int SLOT = kChTiles * RmGeom<T>::CTILE + 2 * kXBlock * 8;  // tiles, X_{j+2}, M1_{j+1}
  static constexpr int BASE = 8192;                                            // control block + rings
};

__device__ __forceinline__ void ch_spin(const volatile int *flag, int need, int32_t *status) {
  if (ld_acquire_cta_smem(flag) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire_cta_smem(flag) < need) {
    if (clock64() - t0 > kWatchdogCycles) { atomicMax(status, CHUB_STATUS_INTERNAL); return; }
  }
}

// LAYOUT: row-major band tiles arrive as two 128-byte-swizzled boxes of [32 rows][16 columns] (lane = row
// reads 16-byte chunks, conflict free); column-major ones as one plain box [32 columns][32 rows] (lane = row
// reads 8 bytes per column, conflict free).
template <typename T, int LAYOUT, bool PROF>
__device__ void df_chain_v4(const DfParams &P, const CUtensorMap *tmap32, unsigned char *smem) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC, EPC = RmGeom<T>::EPC;
  constexpr int CTILE = RmGeom<T>::CTILE, SLOT = ChGeom<T>::SLOT;
  static_assert(kDfD >= 2 && kChTiles == 4, "band warps are dealt block rows modulo 4");
  static_assert(SLOT % 1024 == 0, "ring slots keep the 1024-byte alignment of the swizzled boxes");
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int T_ = P.T;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [kDfSlots]
  volatile int *ctl = reinterpret_cast<volatile int *>(smem + 64);
  volatile int *pdone = ctl;        // p_r is in pbuf for r < *pdone
  volatile int *xwdone = ctl + 1;   // xw_r = X_r wt_r is in xwb for r < *xwdone
  volatile int *xprog = ctl + 7;    // block columns whose X (in the ring slot) warp P has taken into registers
  volatile int *bprog = ctl + 2;    // [4] block columns whose ring slot band warp k has finished reading
  volatile int *uready = ctl + 8;   // [kChRing] r + 1 once u_r is in ubuf[r % kChRing]
  double *pbuf = reinterpret_cast<double *>(smem + 512);   // [kChRing][32]
  double *ubuf = pbuf + kChRing * kPanel;                   // [kChRing][32]
  double *xwb = ubuf + kChRing * kPanel;                    // [kChRing][32]  xw_r
  double *wtb = xwb + kChRing * kPanel;                     // [32] staging of wt_r (warp P)
  double *wsum = wtb + kPanel;                              // [4][32] per band warp: staging of acc_r
  unsigned char *slots = smem + ChGeom<T>::BASE;            // [kDfSlots][SLOT]

  if (tid == 0) {
    for (int i = 0; i < kDfSlots; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  if (tid < 32) ctl[tid] = 0;
  named_sync(1, 256);

  if (warp == 0) {
    // ------------------------------------------------------------ C: the critical path
    long long t_chain0 = DF_PROF_T();
    {
      const double w0 = df_wait_value(P.ws.w + lane, P.status);
      pbuf[lane] = w0;  // (staging: the mat-vec reads all 32 entries)
      __syncwarp();
      const double *X0 = P.ws.X + lane * kXStride;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
      for (int k = 0; k < kPanel; k += 2) {
        a0 = fma(X0[k], pbuf[k], a0);
        a1 = fma(X0[k + 1], pbuf[k + 1], a1);
      }
      __syncwarp();
      const double p0 = a0 + a1;
      pbuf[lane] = p0;
      st_relaxed_f64(P.ws.p + lane, p0);
      __syncwarp();
      if (lane == 0) st_release_cta_smem(pdone, 1);
    }
    // u_1 = X_1 wt_1 (block row 1 has no band tile beyond distance 1)
    double u_first = 0.0;
    if (T_ > 1) {
      const double w1 = df_wait_value(P.ws.w + kPanel + lane, P.status);
      ubuf[kPanel + lane] = w1;  // staging in the slot of u_1
      __syncwarp();
      const double *X1 = P.ws.X + kXBlock + lane * kXStride;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
      for (int k = 0; k < kPanel; k += 2) {
        a0 = fma(X1[k], ubuf[kPanel + k], a0);
        a1 = fma(X1[k + 1], ubuf[kPanel + k + 1], a1);
      }
      u_first = a0 + a1;
      __syncwarp();
    }
    for (int j = 0; j + 1 < T_; ++j) {
      const int r = j + 1;
      if (PROF && P.prof && P.prof_late && j == T_ * 5 / 8) {
        for (int i = 0; i < 8; ++i) pacc[i] = 0;
        t_chain0 = clock64();
      }
      long long tp = DF_PROF_T();
      mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
      if (lane == 0) DF_PROF_ADD(1, tp);
      tp = DF_PROF_T();
      const double *M = reinterpret_cast<const double *>(slots + (size_t)(j % kDfSlots) * SLOT + kChTiles * CTILE +
                                                         kXBlock * 8) + lane * kXStride;
      const double *pj = pbuf + (j % kChRing) * kPanel;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int k = 0; k < kPanel; k += 4) {
        const double2 x = *reinterpret_cast<const double2 *>(M + k);
        const double2 y = *reinterpret_cast<const double2 *>(M + k + 2);
        const double2 a = *reinterpret_cast<const double2 *>(pj + k);
        const double2 b = *reinterpret_cast<const double2 *>(pj + k + 2);
        a0 = fma(x.x, a.x, a0);
        a1 = fma(x.y, a.y, a1);
        a2 = fma(y.x, b.x, a2);
        a3 = fma(y.y, b.y, a3);
      }
      if (lane == 0) DF_PROF_ADD(4, tp);
      tp = DF_PROF_T();
      double u = u_first;
      if (r > 1) {
        ch_spin(uready + (r % kChRing), r + 1, P.status);
        u = ubuf[(r % kChRing) * kPanel + lane];
        if (lane == 0) DF_PROF_ADD(2, tp);
        tp = DF_PROF_T();
        ch_spin(xwdone, r + 1, P.status);
        u += xwb[(r % kChRing) * kPanel + lane];
        if (lane == 0) DF_PROF_ADD(6, tp);
      }
      const double pr = u - ((a0 + a1) + (a2 + a3));
      pbuf[(r % kChRing) * kPanel + lane] = pr;
      st_relaxed_f64(P.ws.p + (int64_t)r * kPanel + lane, pr);  // publish: the data is its own flag
      __syncwarp();
      if (lane == 0) {
        st_release_cta_smem(pdone, r + 1);
        DF_TRACE(2, r);
      }
    }
    if (lane == 0) DF_PROF_ADD(0, t_chain0);
  } else if (warp <= 4) {
    // ------------------------------------------------------------ B: band tiles and u_r
    const int k4 = warp - 1;
    double acc = 0.0;
    double *ws_ = wsum + k4 * kPanel;
    for (int j = 0; j < T_; ++j) {
      if (PROF && P.prof && P.prof_late && j == T_ * 5 / 8) pacc[3] = 0;
      const int r = j + 2 + ((k4 - j) & 3);  // my block row among j+2 .. j+5:  r = k4 + 2 (mod 4)
      if (r < T_) {
        if (j == 0 || r - j == kDfD) acc = 0.0;  // first tile of this block row
        ch_spin(pdone, j + 1, P.status);
        long long tp = DF_PROF_T();
        mbar_wait(&full[j % kDfSlots], (unsigned)((j / kDfSlots) & 1), P.status);
        if (lane == 0 && k4 == 0) DF_PROF_ADD(3, tp);
        const unsigned char *slot = slots + (size_t)(j % kDfSlots) * SLOT;
        const unsigned char *tb = slot + (size_t)(r - j - 2) * 4096;  // [half][tile q][32 rows][128 B]
        const double *pj = pbuf + (j % kChRing) * kPanel;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (LAYOUT == CHUB_COL_MAJOR) {
          // element (tile q, row = lane, column k) at [k][32 q + lane]
          const T *tc = reinterpret_cast<const T *>(slot) + (r - j - 2) * kPanel + lane;
          constexpr int CS = kChTiles * kPanel;
#pragma unroll
          for (int k = 0; k < kPanel; k += 4) {
            const double2 pp = *reinterpret_cast<const double2 *>(pj + k);
            const double2 pq = *reinterpret_cast<const double2 *>(pj + k + 2);
            a0 = fma((double)tc[k * CS], pp.x, a0);
            a1 = fma((double)tc[(k + 1) * CS], pp.y, a1);
            a2 = fma((double)tc[(k + 2) * CS], pq.x, a2);
            a3 = fma((double)tc[(k + 3) * CS], pq.y, a3);
          }
        } else {
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const unsigned char *rowp = tb + h * (kChTiles * 4096) + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int4 raw = *reinterpret_cast<const int4 *>(rowp + ((c ^ (lane & 7)) << 4));
            const T *x = reinterpret_cast<const T *>(&raw);
            const int k0 = h * BOXC + c * EPC;
            if (EPC == 2) {
              const double2 pp = *reinterpret_cast<const double2 *>(pj + k0);
              if (c & 1) { a2 = fma((double)x[0], pp.x, a2); a3 = fma((double)x[1], pp.y, a3); }
              else       { a0 = fma((double)x[0], pp.x, a0); a1 = fma((double)x[1], pp.y, a1); }
            } else {
              const double2 pp = *reinterpret_cast<const double2 *>(pj + k0);
              const double2 pq = *reinterpret_cast<const double2 *>(pj + k0 + 2);
              a0 = fma((double)x[0], pp.x, a0);
              a1 = fma((double)x[1], pp.y, a1);
              a2 = fma((double)x[2 % EPC], pq.x, a2);
              a3 = fma((double)x[3 % EPC], pq.y, a3);
            }
          }
        }
        }
        acc -= (a0 + a1) + (a2 + a3);
        if (r - j == 2) {
          // last band tile of block row r: ua_r = X_r acc; X_r came with this block column
          ws_[lane] = acc;
          __syncwarp();
          const double *X = reinterpret_cast<const double *>(slot + kChTiles * CTILE) + lane * kXStride;
          double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
          for (int k = 0; k < kPanel; k += 4) {
            const double2 x = *reinterpret_cast<const double2 *>(X + k);
            const double2 y = *reinterpret_cast<const double2 *>(X + k + 2);
            const double2 a = *reinterpret_cast<const double2 *>(ws_ + k);
            const double2 b = *reinterpret_cast<const double2 *>(ws_ + k + 2);
            b0 = fma(x.x, a.x, b0);
            b1 = fma(x.y, a.y, b1);
            b2 = fma(y.x, b.x, b2);
            b3 = fma(y.y, b.y, b3);
          }
          ubuf[(r % kChRing) * kPanel + lane] = (b0 + b1) + (b2 + b3);
          __syncwarp();
          if (lane == 0) st_release_cta_smem(uready + (r % kChRing), r + 1);
        }
      }
      __syncwarp();
      if (lane == 0) st_release_cta_smem(bprog + k4, j + 1);  // my reads of ring slot j % kDfSlots are done
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ P: wt_r from the workers (global), xw_r = X_r wt_r
    unsigned long long b0 = kSentinelBits, b1 = kSentinelBits;
    if (T_ > 2) b0 = ld_relaxed_u64(P.ws.w + 2 * kPanel + lane);
    if (T_ > 3) b1 = ld_relaxed_u64(P.ws.w + 3 * kPanel + lane);
    for (int r = 2; r < T_; ++r) {
      if (PROF && P.prof && P.prof_late && r == T_ * 5 / 8) pacc[5] = 0;
      // X_r travels with block column r - 2: my row of it goes into registers, the slot is released
      const int jx = r - 2;
      mbar_wait(&full[jx % kDfSlots], (unsigned)((jx / kDfSlots) & 1), P.status);
      const double *X = reinterpret_cast<const double *>(slots + (size_t)(jx % kDfSlots) * SLOT + kChTiles * CTILE) +
                        lane * kXStride;
      double2 xr[kPanel / 2];
#pragma unroll
      for (int k = 0; k < kPanel / 2; ++k) xr[k] = *reinterpret_cast<const double2 *>(X + 2 * k);
      __syncwarp();
      if (lane == 0) st_release_cta_smem(xprog, jx + 1);
      long long tq = DF_PROF_T();
      const double *nx = r + 1 < T_ ? P.ws.w + (int64_t)(r + 1) * kPanel + lane : nullptr;
      const double wt = df_wait_value2(P.ws.w + (int64_t)r * kPanel + lane, nx, b0, b1, P.status);
      b0 = b1;
      b1 = r + 2 < T_ ? ld_relaxed_u64(P.ws.w + (int64_t)(r + 2) * kPanel + lane) : kSentinelBits;
      if (lane == 0) { DF_PROF_ADD(5, tq); DF_TRACE(1, r); }
      wtb[lane] = wt;
      __syncwarp();
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int k = 0; k < kPanel / 2; k += 2) {
        const double2 a = *reinterpret_cast<const double2 *>(wtb + 2 * k);
        const double2 b = *reinterpret_cast<const double2 *>(wtb + 2 * k + 2);
        a0 = fma(xr[k].x, a.x, a0);
        a1 = fma(xr[k].y, a.y, a1);
        a2 = fma(xr[k + 1].x, b.x, a2);
        a3 = fma(xr[k + 1].y, b.y, a3);
      }
      // xwb[r % 8] was last read for p_{r-8}: C is long past it (p_{r-1} needs ua/xw of r-1 only)
      xwb[(r % kChRing) * kPanel + lane] = (a0 + a1) + (a2 + a3);
      __syncwarp();
      if (lane == 0) st_release_cta_smem(xwdone, r + 1);
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------ T: TMA loads, block column jj -> slot jj % 4
    for (int jj = 0; jj < T_; ++jj) {
      if (kDfPf > 0) {  // band tiles of block column jj + kDfPf -> L2 (nothing to wait for)
        const int jp = jj + kDfPf;
        if (jp + 2 < T_ && elect_one()) {
          if (LAYOUT == CHUB_COL_MAJOR) tma_prefetch_2d(tmap32, (jp + 2) * kPanel, jp * kPanel);
          else tma_prefetch_3d(tmap32, 0, (jp + 2) * kPanel, jp * HALVES);
        }
      }
      if (elect_one()) {
        if (jj >= kDfSlots) {
          // the slot held block column jj - 4: read by the band warps (tiles, X) and by C (M1 -> p_{jj-3})
          const int done = jj - kDfSlots + 1;
          for (int k = 0; k < 4; ++k) ch_spin(bprog + k, done, P.status);
          ch_spin(pdone, min(done + 1, T_), P.status);
          if (done + 1 < T_) ch_spin(xprog, done, P.status);  // (X_{jj-2} travelled with block column jj - 4)
        }
        const int slot = jj % kDfSlots;
        void *bar = &full[slot];
        unsigned char *sb = slots + (size_t)slot * SLOT;
        const bool do_t = jj + 2 < T_, do_m = jj + 1 < T_;  // (do_t: tiles and X_{jj+2})
        const unsigned bytes = (do_t ? (unsigned)kChTiles * CTILE + kXBlock * 8u : 0u) + (do_m ? kXBlock * 8u : 0u);
        if (bytes == 0) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, bytes);
          if (do_m) bulk_g2s(sb + kChTiles * CTILE + kXBlock * 8, P.ws.M1 + (int64_t)(jj + 1) * kXBlock, kXBlock * 8u, bar);
          if (do_t) {  // ONE box: the four band tiles (block rows jj+2 .. jj+5); rows past the end are zero-filled
            if (LAYOUT == CHUB_COL_MAJOR) tma_load_2d(sb, tmap32, (jj + 2) * kPanel, jj * kPanel, bar);
            else tma_load_3d(sb, tmap32, 0, (jj + 2) * kPanel, jj * HALVES, bar);
            bulk_g2s(sb + kChTiles * CTILE, P.ws.X + (int64_t)(jj + 2) * kXBlock, kXBlock * 8u, bar);
          }
        }
      }
      __syncwarp();
    }
  }
  if (P.prof && lane == 0 && (warp == 0 || warp == 1 || warp == 5)) {
    for (int i = 0; i < 8; ++i)
      if (pacc[i]) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
  }
}

template <typename T>
inline size_t df_chain_v4_smem() {
  return (size_t)ChGeom<T>::BASE + (size_t)kDfSlots * ChGeom<T>::SLOT;
}


}  // namespace chub
