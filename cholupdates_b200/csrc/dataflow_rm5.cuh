// dataflow_rm5.cuh -- row warps of the row-major persistent kernel, second generation (included by dataflow.cuh).
//
// What the first generation (df_worker_rm) was bound by (profiles/r02s: with the chain removed and nobody
// waiting for anybody the workers alone take 0.469 ms, the chain CTA alone 0.319 ms): the bottom row warp
// has T + d sequential steps and each one cost ~1700 cycles of mostly fixed latency -- two mbarrier waits,
// four 1 KB TMA loads and four 1 KB TMA stores issued by one lane (~100 cycles each), two dependent
// prefix / scan / update chains, fence, commit.  So the kernel was bound by 517 x 1700 cycles = 0.45 ms of
// one warp's serial latency, not by HBM and not by the chain.
//
// Here a row warp walks its block columns ("pans") in ITERATIONS of two block columns (64 columns):
//   * one 3-D tensor-map TMA instruction moves the whole [8 rows][64 columns] tile (the matrix seen as
//     [column block of 16][row][16 columns]: the box lands as the same 128-byte-swizzled 1 KB boxes the
//     2-D path produced, so the lane mapping and bank behaviour are unchanged) -- 2 TMA ops per iteration
//     instead of 8; one mbarrier per iteration;
//   * the two tiles' prefix products, scans and updates are written as one straight-line block, so the
//     two dependency chains overlap;
//   * far block columns (pan < b - d) are gated on their own p; the band columns (b-d .. b) are gated on
//     p_b alone (the chain has consumed the old band entries of block row b once p_b exists), no idle steps;
//   * the residual hand-over to the chain happens right after the prefix sums of the last far pair.
// The coefficient ring is recycled by pans (prog[] = next pan a warp still needs) instead of steps.
#pragma once

namespace chub {

// ---- relay warp: p_s from global memory (the chain's NaN-sentinel publication) into the coefficient ring.
// Split from the coefficient warp: a row warp's far iterations need p only, and one warp doing poll + scan +
// rsqrt per block column (~1100 cycles) was what the row warps waited for (profiles/r02v).
template <bool PROF>
__device__ __forceinline__ void df_relay_warp(const DfParams &P, double *coef, volatile int *prog, volatile int *pprog,
                                              int wkr, int lane, int reuse_lag = 1) {
  const int T_ = P.T;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long b0 = ld_relaxed_u64(P.ws.p + lane);
  for (int pan = 0; pan < T_; ++pan) {
    const int k = pan * kPanel + lane;
    long long tq = DF_PROF_T();
    double pk;
#ifdef CHUB_EXP_WORKER_FREE  // timing experiment (results are garbage): the workers never wait for the chain
    pk = b0 != kSentinelBits ? __longlong_as_double((long long)b0) : 0.0;
#else
    pk = b0 != kSentinelBits ? __longlong_as_double((long long)b0) : df_wait_value(P.ws.p + k, P.status);
#endif
    b0 = pan + 1 < T_ ? ld_relaxed_u64(P.ws.p + k + kPanel) : kSentinelBits;
    if (lane == 0) DF_PROF_ADD(6, tq);
    if (lane == 0 && wkr == 0) DF_TRACE(3, pan);
    if (pan >= kDfCoefRing) {  // the slot's previous block column (pan - ring) is no longer needed by any row warp
      // prog[] = the next block column a row warp still needs (rm5, reuse_lag = 1) or the steps it has finished
      // (column-major workers: block column pan - ring is read for the last time at step pan - ring + d)
      const int need = pan - kDfCoefRing + reuse_lag;
      const long long t0 = clock64();
      while (__reduce_min_sync(0xffffffffu, ld_acquire_cta_smem(prog + lane)) < need) {
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
      }
    }
    coef[(size_t)(pan % kDfCoefRing) * 4 * kPanel + lane] = pk;
    __syncwarp();
    if (lane == 0) st_release_cta_smem(pprog, pan + 1);
    if (lane == 0) DF_PROF_ADD(7, tq);
  }
  if (PROF && P.prof && lane == 0) {
    P.prof[(size_t)blockIdx.x * 16 + 6] = pacc[6];
    P.prof[(size_t)blockIdx.x * 16 + 7] = pacc[7];
  }
}

// ---- coefficient warp: (c, sigma, L'_kk) of two block columns per iteration from the p the relay warp put
// into the ring; the two scans / rsqrt chains are independent up to one addition (t after the first column)
template <typename T, bool UPDATE, bool PROF>
__device__ __forceinline__ void df_coef_warp5(T *v, const DfParams &P, double *coef, volatile int *cprog,
                                              volatile int *pprog, int wkr, int lane) {
  if (!UPDATE) return;  // forward substitution only: the row warps never look at cprog
  const int n = P.n, T_ = P.T, Wk = P.Wk;
  double t_run = P.ws.scal[0], S_run = P.ws.scal[1];  // (1, 1) unless a blocked driver carries them over
  // the input diagonal of the next three block columns, prefetched (an L2 round trip each)
  double d0 = P.ws.dg[lane], d1 = T_ > 1 ? P.ws.dg[kPanel + lane] : 1.0, d2 = T_ > 2 ? P.ws.dg[2 * kPanel + lane] : 1.0;
  for (int pan = 0; pan < T_;) {
    if (ld_acquire_cta_smem(pprog) < pan + 1) {
      const long long t0 = clock64();
      while (ld_acquire_cta_smem(pprog) < pan + 1) {
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
      }
    }
    // two block columns at once when the second one's p is already there (the row-major row warps consume pairs,
    // and while the chain runs ahead of the workers it always is); never wait for it (the column-major row
    // warps take one block column per step: waiting would add a chain period to every other step)
    const bool two = pan + 1 < T_ && ld_acquire_cta_smem(pprog) >= pan + 2;
    const int upto = pan + (two ? 2 : 1);
    double *cfA = coef + (size_t)(pan % kDfCoefRing) * 4 * kPanel;
    double *cfB = coef + (size_t)((pan + 1) % kDfCoefRing) * 4 * kPanel;
    const double pA = cfA[lane], pB = two ? cfB[lane] : 0.0;
    const double dgA = d0, dgB = two ? d1 : 1.0;
    if (two) {
      d0 = d2;
      d1 = pan + 3 < T_ ? P.ws.dg[(pan + 3) * kPanel + lane] : 1.0;
      d2 = pan + 4 < T_ ? P.ws.dg[(pan + 4) * kPanel + lane] : 1.0;
    } else {
      d0 = d1;
      d1 = d2;
      d2 = pan + 3 < T_ ? P.ws.dg[(pan + 3) * kPanel + lane] : 1.0;
    }
    double inA = pA * pA, inB = pB * pB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double ya = __shfl_up_sync(0xffffffffu, inA, o);
      const double yb = __shfl_up_sync(0xffffffffu, inB, o);
      if (lane >= o) { inA += ya; inB += yb; }
    }
    double exA = __shfl_up_sync(0xffffffffu, inA, 1), exB = __shfl_up_sync(0xffffffffu, inB, 1);
    if (lane == 0) { exA = 0.0; exB = 0.0; }
    const double t_mid = t_run + __shfl_sync(0xffffffffu, inA, 31);
    const double tkA = t_run + exA, tk1A = t_run + inA;
    const double tkB = t_mid + exB, tk1B = t_mid + inB;
    const unsigned negA = __ballot_sync(0xffffffffu, dgA < 0.0), negB = __ballot_sync(0xffffffffu, dgB < 0.0);
    const double sgnA = dgA < 0.0 ? -1.0 : 1.0, sgnB = dgB < 0.0 ? -1.0 : 1.0;
    const double rinvA = rsqrt(tkA * tk1A), rinvB = rsqrt(tkB * tk1B);
    const double ckA = sgnA * tkA * rinvA, sgA = sgnA * pA * rinvA;
    const double ckB = sgnB * tkB * rinvB, sgB = sgnB * pB * rinvB;
    cfA[kPanel + lane] = ckA;
    cfA[2 * kPanel + lane] = sgA;
    cfA[3 * kPanel + lane] = fabs(dgA) * tk1A * rinvA;
    if (two) {
      cfB[kPanel + lane] = ckB;
      cfB[2 * kPanel + lane] = sgB;
      cfB[3 * kPanel + lane] = fabs(dgB) * tk1B * rinvB;
    }
    __syncwarp();
    if (lane == 0) st_release_cta_smem(cprog, upto);
    if (lane == 0 && wkr == 0) DF_TRACE(4, pan);
    // one worker per block column: coefficients for the rectangles below a diagonal block, rotg's z_k in v[k]
    if (wkr == pan % Wk) {
      const int k = pan * kPanel + lane;
      P.ws.c[k] = ckA;
      P.ws.sg[k] = sgA;
      if (k < n) {
        const double Sk = (__popc(negA & ((1u << lane) - 1u)) & 1) ? -S_run : S_run;
        const double nuk = Sk * dgA * pA / sqrt(tkA);
        const double sk = Sk * sgnA * pA / sqrt(tk1A);
        v[k] = (T)rotg_z<double>(dgA, nuk, ckA, sk);
      }
    }
    if (__popc(negA) & 1) S_run = -S_run;
    if (two && wkr == (pan + 1) % Wk) {
      const int k = (pan + 1) * kPanel + lane;
      P.ws.c[k] = ckB;
      P.ws.sg[k] = sgB;
      if (k < n) {
        const double Sk = (__popc(negB & ((1u << lane) - 1u)) & 1) ? -S_run : S_run;
        const double nuk = Sk * dgB * pB / sqrt(tkB);
        const double sk = Sk * sgnB * pB / sqrt(tk1B);
        v[k] = (T)rotg_z<double>(dgB, nuk, ckB, sk);
      }
    }
    if (two && (__popc(negB) & 1)) S_run = -S_run;
    t_run = two ? t_mid + __shfl_sync(0xffffffffu, inB, 31) : t_mid;
    pan = upto;
  }
  if (wkr == 0 && lane == 0) { P.ws.scal[4] = t_run; P.ws.scal[5] = S_run; }
}

constexpr int kRm5Slots = 3;  // pair slots per row warp (each = two stages of the old ring: same shared memory)

template <typename T, bool UPDATE, bool PROF>
__device__ void df_worker_rm5(T *L, T *v, const DfParams &P, const CUtensorMap *tmap8, const CUtensorMap *tmap3,
                              unsigned char *smem) {
  constexpr int HALVES = RmGeom<T>::HALVES, BOXC = RmGeom<T>::BOXC;
  constexpr int CPT = RmGeom<T>::CPT, WTILE = RmGeom<T>::WTILE;
  static_assert(kDfStages == 2 * kRm5Slots, "pair slots reuse the six-stage ring");
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int wkr = blockIdx.x - 1;
  const int n = P.n, T_ = P.T, Wk = P.Wk, nrw = P.nw;
  const int64_t ld = P.ld;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);  // [nrw][kDfStages] (kRm5Slots used)
  unsigned long long *cfull = full + kDfStages * 2 * kRmMaxQ;               // (unused here; keeps the layout)
  volatile int *prog = reinterpret_cast<volatile int *>(cfull + kDfCoefRing);  // [32] next pan each row warp needs
  volatile int *cprog = prog + 32;  // block columns whose coefficients (c, sigma, dn) are in the ring
  volatile int *pprog = prog + 33;  // block columns whose p is in the ring (always >= *cprog)
  double *coef = reinterpret_cast<double *>(smem + 2048);                   // [kDfCoefRing][4][32]
  unsigned char *stages = smem + 2048 + kDfCoefRing * 4 * kPanel * 8;       // 1024-aligned

  if (tid == 0) {
    for (int i = 0; i < kDfStages * nrw; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  if (tid < 32) prog[tid] = tid < nrw ? 0 : 0x7fffffff;
  if (tid == 0) { *cprog = 0; *pprog = 0; }
  __syncthreads();
  if (warp > nrw + 1) return;

  if (warp == nrw + 1) {
    df_relay_warp<PROF>(P, coef, prog, pprog, wkr, lane);
    return;
  }
  if (warp == nrw) {
    df_coef_warp5<T, UPDATE, PROF>(v, P, coef, cprog, pprog, wkr, lane);
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
  const int64_t row = row0 + r8;
  double w = row < n ? P.ws.v0[row] : 0.0;
  const int64_t row0_next = (((int64_t)wkr + (int64_t)((warp + 1) >> 1) * Wk) << 4) + (((warp + 1) & 1) << 3);
  const bool prof_me = PROF && P.prof && lane == 0 && (warp + 1 >= nrw || row0_next >= n);
  // my 8 columns live in half `hh`, chunks cb .. cb+CPT-1 of the 128-byte box row
  const int hh = (g * 8 * (int)sizeof(T)) / 128;
  const int cb = ((g * 8 * (int)sizeof(T)) % 128) / 16;
  unsigned long long *myfull = full + warp * kDfStages;
  unsigned char *mytile = stages + (size_t)warp * (kDfStages * WTILE);  // [kRm5Slots][2][WTILE]
  const unsigned my_off = hh * 1024 + r8 * 128;
  const int n3 = P.use3d ? (n / BOXC) * BOXC : 0;  // columns the 3-D view covers

  // ---- the warp's schedule: far pans 0 .. F-1, band pans F .. b-1 (full tiles), the diagonal tile b
  const int F = max(0, b - kDfD);
  const int fs = F & 1, nfar = fs + (F >> 1);
  const int Bf = b - F, bs = Bf & 1, nbf = bs + (Bf >> 1);
  const int NI = nfar + nbf + 1;
  auto iter = [&](int it, int &pan0, int &cnt) {
    if (it < nfar) {
      if (fs) { pan0 = it == 0 ? 0 : 2 * it - 1; cnt = it == 0 ? 1 : 2; }
      else { pan0 = 2 * it; cnt = 2; }
    } else {
      const int j = it - nfar;
      if (j < nbf) {
        if (bs) { pan0 = j == 0 ? F : F + 2 * j - 1; cnt = j == 0 ? 1 : 2; }
        else { pan0 = F + 2 * j; cnt = 2; }
      } else { pan0 = b; cnt = 1; }
    }
  };

  auto issue_load = [&](int it) {  // elected lane only
    if (it >= NI) return;
    int pan0, cnt;
    iter(it, pan0, cnt);
    const int slot = it % kRm5Slots;
    void *bar = &myfull[slot];
    unsigned char *tile = mytile + slot * 2 * WTILE;
    const int c0 = pan0 * kPanel;
    if (cnt == 2 && c0 + 2 * kPanel <= n3) {
      mbar_arrive_expect_tx(bar, 2u * WTILE);
      tma_load_3d(tile, tmap3, 0, (int)row0, c0 / BOXC, bar);
      return;
    }
    unsigned bytes = 0;
    for (int t = 0; t < cnt; ++t) bytes += (HALVES == 2 && c0 + t * kPanel + BOXC < n) ? 2048u : 1024u;
    mbar_arrive_expect_tx(bar, bytes);
    for (int t = 0; t < cnt; ++t) {
      const int cc = c0 + t * kPanel;
      tma_load_2d(tile + t * WTILE, tmap8, cc, (int)row0, bar);
      if (HALVES == 2 && cc + BOXC < n) tma_load_2d(tile + t * WTILE + 1024, tmap8, cc + BOXC, (int)row0, bar);
    }
  };
  auto issue_store = [&](int pan0, int cnt, int slot) {  // elected lane only; full tiles only
    unsigned char *tile = mytile + slot * 2 * WTILE;
    const int c0 = pan0 * kPanel;
    if (cnt == 2 && c0 + 2 * kPanel <= n3) {
      tma_store_3d(tmap3, 0, (int)row0, c0 / BOXC, tile);
      return;
    }
    for (int t = 0; t < cnt; ++t) {
      const int cc = c0 + t * kPanel;
      tma_store_2d(tmap8, cc, (int)row0, tile + t * WTILE);
      if (HALVES == 2 && cc + BOXC < n) tma_store_2d(tmap8, cc + BOXC, (int)row0, tile + t * WTILE + 1024);
    }
  };

  // ---- arithmetic of one tile, in three pieces so that a pair can interleave them
  // (1) load my 8 entries and p, local prefix products: pre[j] = sum_{m<j} lo_m p_m, run = the segment total
  auto tile_prefix = [&](unsigned char *rowp, const double *cf, int4 (&raw)[CPT], double (&lo)[8], double (&pre)[8],
                         double &run, bool diag, int64_t c0) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) raw[i] = *reinterpret_cast<const int4 *>(rowp + (((cb + i) ^ r8) << 4));
    const T *x = reinterpret_cast<const T *>(raw);
    double pc[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 a = *reinterpret_cast<const double2 *>(cf + j);
      pc[j] = a.x; pc[j + 1] = a.y;
    }
    run = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      lo[j] = (double)x[j];
      pre[j] = run;
      if (!diag || c0 + j < row) run = fma(lo[j], pc[j], run);
    }
  };
  // (2) exclusive scan of the four segment totals of a row (lanes g*8 + r8, same r8)
  auto seg_scan = [&](double run, double &excl, double &total) {
    double incl = run;
    double y = __shfl_up_sync(0xffffffffu, incl, 8);
    if (g >= 1) incl += y;
    y = __shfl_up_sync(0xffffffffu, incl, 16);
    if (g >= 2) incl += y;
    total = __shfl_sync(0xffffffffu, incl, 24 + r8);
    excl = incl - run;
  };
  // (3) L_ik <- c_k L_ik + sigma_k (w_i - sum_{m<k} L_im p_m), back into the tile
  auto tile_update = [&](unsigned char *rowp, const double *cf, int4 (&raw)[CPT], const double (&lo)[8],
                         const double (&pre)[8], double wseg) {
    T *x = reinterpret_cast<T *>(raw);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 c = *reinterpret_cast<const double2 *>(cf + kPanel + j);
      const double2 d = *reinterpret_cast<const double2 *>(cf + 2 * kPanel + j);
      x[j] = (T)fma(c.x, lo[j], d.x * (wseg - pre[j]));
      x[j + 1] = (T)fma(c.y, lo[j + 1], d.y * (wseg - pre[j + 1]));
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) *reinterpret_cast<int4 *>(rowp + (((cb + i) ^ r8) << 4)) = raw[i];
  };
  auto wait_count = [&](volatile int *ctr, int need) {
    if (ld_acquire_cta_smem(ctr) < need) {
      const long long t0 = clock64();
      while (ld_acquire_cta_smem(ctr) < need) {
        if (clock64() - t0 > kWatchdogCycles) { atomicMax(P.status, CHUB_STATUS_INTERNAL); break; }
      }
    }
  };

  if (elect_one()) {
    issue_load(0);
    issue_load(1);
  }
  __syncwarp();
  const long long t_wk0 = DF_PROF_T();
  for (int it = 0; it < NI; ++it) {
    long long tp = DF_PROF_T();
    int pan0, cnt;
    iter(it, pan0, cnt);
    const int slot = it % kRm5Slots;
    const bool far = it < nfar;
    const bool diag = it == NI - 1;
    // the hand-over iteration (last far pair) takes its two block columns as their p arrive: it is on the
    // chain <-> worker latency loop, and the first column's p is there one chain period before the second's
    const bool handover = it == nfar - 1;
    const int gate = far ? (handover ? pan0 : pan0 + cnt - 1) : b;
    mbar_wait(&myfull[slot], (unsigned)((it / kRm5Slots) & 1), P.status);
    if (prof_me) DF_PROF_ADD(1, tp);
    tp = DF_PROF_T();
    wait_count(pprog, gate + 1);
    if (prof_me) DF_PROF_ADD(4, tp);
    if (prof_me && wkr == 0 && far) DF_TRACE(5, pan0);
    tp = DF_PROF_T();

    unsigned char *rowA = mytile + slot * 2 * WTILE + my_off;
    const double *cfA = coef + (size_t)(pan0 % kDfCoefRing) * 4 * kPanel + g * 8;
    const int64_t c0A = (int64_t)pan0 * kPanel + g * 8;  // my first column
    bool store = false;
    if (cnt == 2) {
      unsigned char *rowB = rowA + WTILE;
      const double *cfB = coef + (size_t)((pan0 + 1) % kDfCoefRing) * 4 * kPanel + g * 8;
      int4 rawA[CPT], rawB[CPT];
      double loA[8], preA[8], loB[8], preB[8], runA, runB, exA, exB, totA, totB;
      tile_prefix(rowA, cfA, rawA, loA, preA, runA, false, c0A);
      if (handover) {
        seg_scan(runA, exA, totA);
        wait_count(pprog, pan0 + 2);
      }
      tile_prefix(rowB, cfB, rawB, loB, preB, runB, false, c0A + kPanel);
      if (!handover) seg_scan(runA, exA, totA);
      seg_scan(runB, exB, totB);
      const double wsegA = w - exA;
      w -= totA;
      const double wsegB = w - exB;
      w -= totB;
      // hand the residual over to the chain after the last far block column -- as early as possible: this
      // store is on the chain <-> worker latency loop
      if (handover && g == 0 && row < n) {
        st_relaxed_f64(P.ws.w + row, df_sanitize(w));
        if (lane == 0 && (row0 & 31) == 0) DF_TRACE(0, b);
      }
      if (UPDATE) {
        wait_count(cprog, pan0 + 2);
        tile_update(rowA, cfA, rawA, loA, preA, wsegA);
        tile_update(rowB, cfB, rawB, loB, preB, wsegB);
        store = true;
      }
    } else {
      int4 rawA[CPT];
      double loA[8], preA[8], runA, exA, totA;
      tile_prefix(rowA, cfA, rawA, loA, preA, runA, diag, c0A);
      seg_scan(runA, exA, totA);
      const double wsegA = w - exA;
      w -= totA;
      if (handover && g == 0 && row < n) {
        st_relaxed_f64(P.ws.w + row, df_sanitize(w));
        if (lane == 0 && (row0 & 31) == 0) DF_TRACE(0, b);
      }
      if (UPDATE) {
        wait_count(cprog, pan0 + 1);
        if (!diag) {
          tile_update(rowA, cfA, rawA, loA, preA, wsegA);
          store = true;
        } else if (row < n) {  // the diagonal tile is ragged: written straight to global memory
          T *gp = L + row * ld + c0A;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int64_t col = c0A + j;
            if (col < row) gp[j] = (T)fma(cfA[kPanel + j], loA[j], cfA[2 * kPanel + j] * (wsegA - preA[j]));
            else if (col == row) gp[j] = (T)cfA[3 * kPanel + j];
          }
        }
      }
    }
    if (prof_me) DF_PROF_ADD(3, tp);
    if (prof_me && wkr == 0 && far) DF_TRACE(6, pan0);
    tp = DF_PROF_T();
    // The ring holds the iteration being processed and two iterations of loads in flight.  Iteration it+2
    // reuses the slot of iteration it-1, whose TMA store has had this whole compute phase to drain.
    if (elect_one()) {
      if (it >= 1) bulk_wait_read<0>();
      issue_load(it + 2);
    }
    if (prof_me) DF_PROF_ADD(2, tp);
    tp = DF_PROF_T();
    if (store) fence_proxy_async();  // my generic-proxy writes -> visible to the TMA (async proxy) store
    __syncwarp();
    if (elect_one()) {
      // progress for the coefficient warp (ring-slot reuse): the next pan this warp needs.  Every lane's
      // coefficient reads of this iteration are complete (their values went into the tile writes the fence
      // above ordered, and the __syncwarp orders the lanes), so a plain store is enough.
      prog[warp] = diag ? 0x7fffffff : pan0 + cnt;
      if (store) {
        issue_store(pan0, cnt, slot);
        bulk_commit();
      }
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

}  // namespace chub
