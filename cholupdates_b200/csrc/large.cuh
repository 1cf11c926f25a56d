// large.cuh -- multi-CTA kernels for one large factor ("panel path").
//
// Algorithm (Gill-Golub-Murray-Saunders closed form of the
// Seeger / LINPACK dchud rotations).  With p = L^{-1} v, t_0 = 1, t_{k+1} = t_k + p_k^2
// and the running forward-substitution residual  w_i^{(k)} = v_i - sum_{m<k} L_im p_m :
//
//     c_k      = sgn_k sqrt(t_k / t_{k+1})            (sgn_k = sign of the input L_kk)
//     sigma_k  = sgn_k p_k / sqrt(t_k t_{k+1})
//     L'_ik    = c_k L_ik + sigma_k w_i^{(k)}          (i > k)   == c_k L_ik + s_k nu_i
//     L'_kk    = |L_kk| sqrt(t_{k+1} / t_k)                      == hypot(L_kk, nu_k)
//     w_i^{(k+1)} = w_i^{(k)} - L_ik p_k
//
// which is algebraically the reference's loop (_seeger_impl_cython.pyx:92-155:
// rotg on (L_kk, v_k), rot on (L[k+1:,k], v[k+1:])) with nu_i^{(k)} = S_k w_i^{(k)} / sqrt(t_k),
// S_k = prod_{m<k} sgn_m.  The only sequential part is the blocked forward substitution
// for p (the "chain"): per 32-column block one 32x32 mat-vec with the pre-inverted diagonal
// block, no sqrt / div on the chain.  Everything else is row-independent.
//
// The downdate (.pyx:160-330) reuses the same chain for p = L^{-1} v (read-only pass),
// then rho^2 = 1 - p.p, q_k^2 = rho^2 + sum_{j>=k} p_j^2, c_k = q_{k+1}/q_k, s_k = p_k/q_k and
// the row-independent reverse sweep of appendix A.3.
//
// This file is the launch-per-panel version: no inter-CTA waiting, works for every n and
// both layouts.  The persistent dataflow kernel (dataflow.cuh) reuses its pieces.
#pragma once

#ifndef CHUB_CHAIN_V7
#define CHUB_CHAIN_V7 0  // 1: chain CTA with one mat-vec per block row (dataflow_chain_v7.cuh, needs M1 from the pre-pass); 0: v6
#endif

#include <cstdlib>

#include "common.cuh"

namespace chub {

constexpr int kBig = 256;  // columns per chain-kernel launch (8 blocks of 32)
constexpr int kXStride = kPanel + 2;         // padded row of an inverse diagonal block (doubles):
constexpr int kXBlock = kPanel * kXStride;  // lane = row reads of 16 bytes are bank-conflict free

struct LargeWs {
  double *X;      // [nb][kXBlock] inverse of each 32x32 diagonal block (rows of kXStride, identity padded)
  double *M1;     // [nb][2][kXBlock] M1_r = X_r L[r, r-1], M2_r = X_r L[r, r-2] (dataflow kernel: the chain's one mat-vec
                  // per block row on its critical path, dataflow_chain_v7.cuh)
  double *w;      // [n_pad]  forward-substitution residual, initialised to v
  double *p;      // [n_pad]  p = L^{-1} v
  double *c;      // [n_pad]  c_k (sign folded)
  double *sg;     // [n_pad]  update: sigma_k (sign folded); downdate: s_k
  double *dn;     // [n_pad]  update: new diagonal L'_kk;     downdate: sign of the input diagonal
  double *v0;     // [n_pad]  dataflow kernel: pristine copy of v
  double *dg;     // [n_pad]  dataflow kernel: the input diagonal
  double *scal;   // [8]      0: t (running), 1: S (running sign), 2: p.p, 3: rho^2
  int32_t *flags; // [..]     spare
};

constexpr int64_t kWsPadMax = 8192;  // doubles

// Where the polled vectors sit relative to each other matters: the persistent kernel's chain CTA polls w
// (handed over by the workers) while 147 coefficient warps poll p, and a layout that puts those lines on
// the same L2 slice costs 3.5 % of the n = 16384 update (measured by accident: inserting a 4.4 MB array
// before w took 0.466 -> 0.482 ms with a byte-identical kernel).  CHUB_WS_PAD_W / CHUB_WS_PAD_P (doubles,
// multiples of 32) shift w and p for such experiments; the defaults are the measured best.
inline int64_t ws_pad(const char *name, int64_t dflt) {
  const char *e = getenv(name);
  int64_t x = e ? atoll(e) : dflt;
  if (x < 0) x = 0;
  if (x > kWsPadMax) x = kWsPadMax;
  return x & ~31LL;
}

inline size_t large_ws_doubles(int64_t n) {
  int64_t n_pad = (n + 31) & ~31LL;
  int64_t nb = n_pad / 32;
  // (+ 2 * kWsPadMax: room for the two tuning pads of large_ws_carve)
  return (size_t)(nb * kXBlock + 7 * n_pad + 8 + 2 * kWsPadMax
#if CHUB_CHAIN_V7
                  + 2 * nb * kXBlock
#elif defined(CHUB_CHAIN_V4)
                  + nb * kXBlock
#endif
  );
}

inline LargeWs large_ws_carve(void *base, int64_t n) {
  int64_t n_pad = (n + 31) & ~31LL;
  int64_t nb = n_pad / 32;
  LargeWs ws;
  double *d = reinterpret_cast<double *>(base);
  static const int64_t pad_w = ws_pad("CHUB_WS_PAD_W", 0), pad_p = ws_pad("CHUB_WS_PAD_P", 0);
  ws.X = d; d += nb * kXBlock;
  d += pad_w;
  ws.w = d; d += n_pad;
  d += pad_p;
  ws.p = d; d += n_pad;
  ws.c = d; d += n_pad;
  ws.sg = d; d += n_pad;
  ws.dn = d; d += n_pad;
  ws.v0 = d; d += n_pad;
  ws.dg = d; d += n_pad;
  ws.scal = d; d += 8;
  d += 2 * kWsPadMax - pad_w - pad_p;  // (the flags keep their place whatever the pads)
#if CHUB_CHAIN_V7
  ws.M1 = d; d += 2 * nb * kXBlock;  // [nb][M1 | M2]
#elif defined(CHUB_CHAIN_V4)
  ws.M1 = d; d += nb * kXBlock;
#else
  ws.M1 = nullptr;
#endif
  ws.flags = reinterpret_cast<int32_t *>(d);
  return ws;
}

// ---------------------------------------------------------------- prepare
// check_diag (_arg_validation.py:18-22), w <- v, t <- 1, S <- 1.
template <typename T, int LAYOUT>
__global__ void large_prepare_kernel(const T *L, int n, int64_t ld, const T *v, LargeWs ws,
                                     int flags, int32_t *status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    ws.w[i] = (double)v[i];
    if ((flags & CHUB_FLAG_CHECK_DIAG) && L[idx<LAYOUT>(i, i, ld)] == T(0))
      atomicMax(status, CHUB_STATUS_ZERO_DIAG);
  }
  if (i == 0) {
    ws.scal[0] = 1.0; ws.scal[1] = 1.0; ws.scal[2] = 0.0; ws.scal[3] = 0.0;
  }
}

// --------------------------------------------------------- diagonal inverses
// One warp per 32x32 diagonal block: X = inv(L_D) (lower triangular), lane j computes column j
// by forward substitution (reciprocal diagonal precomputed, two accumulators per dot product so
// that the dependent FMA chain is half as long).  Blocks that stick out of the matrix are padded
// with the identity.  Must be called by all 4 warps of a 128-thread CTA (static shared memory).
// Core: one warp, staging area `Lw` = [32][33] doubles + `Rw` = [32] doubles of shared memory.
template <typename T, int LAYOUT>
__device__ __forceinline__ void invdiag_block_smem(const T *L, int n, int64_t ld, double *X, int blk,
                                                   double (*Lw)[kPanel + 1], double *Rw, double *M1 = nullptr);
constexpr int kInvStageDoubles = kPanel * (kPanel + 1) + kPanel;  // per warp

template <typename T, int LAYOUT>
__device__ __forceinline__ void invdiag_block_to(const T *L, int n, int64_t ld, double *X, int blk,
                                                 double *M1 = nullptr) {
  __shared__ double Ls[4][kPanel][kPanel + 1];
  __shared__ double Rd[4][kPanel];
  const int warp = (threadIdx.x >> 5) & 3;
  invdiag_block_smem<T, LAYOUT>(L, n, ld, X, blk, Ls[warp], Rd[warp], M1);
}

// M1 != nullptr: also M1 = X L[blk, blk-1] (32 x 32, rows of kXStride; zero for blk == 0): with it the
// chain of the persistent kernel needs ONE mat-vec per block row on its critical path,
//     p_r = X_r (wt_r - sum_{q >= 2} L[r, r-q] p_{r-q}) - M1_r p_{r-1}.
template <typename T, int LAYOUT>
__device__ __forceinline__ void invdiag_block_smem(const T *L, int n, int64_t ld, double *X, int blk,
                                                   double (*Lw)[kPanel + 1], double *Rw, double *M1) {
  const int lane = threadIdx.x & 31;
  const int nb = (n + 31) / 32;
  if (blk >= nb) return;
  const int c0 = blk * kPanel;
  // stage the block (lane = column for row-major reads, lane = row for column-major reads).  All 32 loads are
  // issued before the first one is used: one memory round trip instead of eight (the loop as the compiler
  // unrolled it, four loads at a time, was about half of the pre-pass's 8 us).
  {
    double stg[kPanel];
#pragma unroll
    for (int r = 0; r < kPanel; ++r) {
      int i, j;
      if (LAYOUT == CHUB_ROW_MAJOR) { i = r; j = lane; } else { i = lane; j = r; }
      stg[r] = (i == j) ? 1.0 : 0.0;
      if (c0 + i < n && c0 + j < n && j <= i) stg[r] = (double)L[idx<LAYOUT>(c0 + i, c0 + j, ld)];
    }
#pragma unroll
    for (int r = 0; r < kPanel; ++r) {
      if (LAYOUT == CHUB_ROW_MAJOR) Lw[r][lane] = stg[r];
      else Lw[lane][r] = stg[r];
    }
  }
  __syncwarp();
  Rw[lane] = 1.0 / Lw[lane][lane];
  __syncwarp();
  double x[kPanel];
  const int j = lane;
#pragma unroll
  for (int i = 0; i < kPanel; ++i) {
    double a0 = (i == j) ? 1.0 : 0.0, a1 = 0.0;
#pragma unroll
    for (int m = 0; m < kPanel; m += 2) {
      if (m < i) a0 -= Lw[i][m] * x[m];          // x[m] = 0 for m < j
      if (m + 1 < i) a1 -= Lw[i][m + 1] * x[m + 1];
    }
    x[i] = (i >= j) ? (a0 + a1) * Rw[i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < kPanel; ++i) X[i * kXStride + j] = x[i];
#if defined(CHUB_CHAIN_V4) || CHUB_CHAIN_V7  // (chains that do not use M1: compiled out, the dead code alone made the
                                             //  pre-pass measurably slower)
  if (M1 == nullptr) return;
  // M_q[i][c] = sum_{k <= i} X[i][k] Lt[k][c],  Lt = L[blk, blk-q], q = 1, 2;  lane = column c.  X goes through the
  // staging area (its L_D copy is no longer needed), Lt[.][c] sits in registers.
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kPanel; ++i) Lw[i][j] = x[i];
  __syncwarp();
#if CHUB_CHAIN_V7
  constexpr int NQ = 2;
#else
  constexpr int NQ = 1;
#endif
  for (int q = 1; q <= NQ; ++q) {
    double lt[kPanel];
#pragma unroll
    for (int k = 0; k < kPanel; ++k)
      lt[k] = (blk >= q && c0 + k < n) ? (double)L[idx<LAYOUT>(c0 + k, c0 - q * kPanel + lane, ld)] : 0.0;
    double *Mq = M1 + (q - 1) * kXBlock;
#pragma unroll
    for (int i = 0; i < kPanel; ++i) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int k = 0; k < kPanel; k += 2) {
        if (k <= i) a0 = fma(Lw[i][k], lt[k], a0);
        if (k + 1 <= i) a1 = fma(Lw[i][k + 1], lt[k + 1], a1);
      }
      Mq[i * kXStride + lane] = a0 + a1;
    }
  }
#else
  (void)M1;
#endif
}
template <typename T, int LAYOUT>
__device__ __forceinline__ void invdiag_block(const T *L, int n, int64_t ld, LargeWs ws, int blk, bool with_m1 = false) {
  invdiag_block_to<T, LAYOUT>(L, n, ld, ws.X + (int64_t)blk * kXBlock, blk,
                              with_m1 ? ws.M1 + (int64_t)blk * (CHUB_CHAIN_V7 ? 2 : 1) * kXBlock : nullptr);
}

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(128) large_invdiag_kernel(const T *L, int n, int64_t ld,
                                                             LargeWs ws, const int32_t *status,
                                                             int blk0 = 0, int nblk = 1 << 30) {
  if (*status != 0) return;
  const int blk = blk0 + blockIdx.x * 4 + (threadIdx.x >> 5);
  if (blk >= blk0 + nblk) return;
  invdiag_block<T, LAYOUT>(L, n, ld, ws, blk);
}

// ------------------------------------------------------------------ chain
// One CTA, 256 threads, one launch per 256 columns.  Thread a <-> row J*256 + a.
// For each of the 8 diagonal blocks: p_blk = X_blk w_blk (mat-vec, 8 threads per row), then the
// remaining rows of the 256-row block subtract L[:, blk] p_blk from their w (kept in a register).
// UPDATE: afterwards the Seeger coefficients of the 256 columns are formed from a prefix sum of
// p^2 and written to the workspace; v[k] <- rotg's z_k.
//
// chain_body is the kernel's body with every buffer explicit, so that the wavefront row-cyclic mode
// (rowcyclic.cuh) can point it at a task's payload: Xp = the 8 inverse diagonal blocks of panel J,
// w_in indexed by the global row, p/c/sg/dn/z outputs indexed by a = column - J*256 (z_out and v may
// each be null), scal_in = {t, S} entering the panel, scal_out = {t, S} leaving it.
template <typename T, int LAYOUT, bool UPDATE>
__device__ __forceinline__ void chain_body(const T *L, int n, int64_t ld, int J, const double *Xp,
                                           const double *w_in, double *p_out, double *c_out,
                                           double *sg_out, double *dn_out, T *v, double *z_out,
                                           const double *scal_in, double *scal_out) {
  __shared__ double wsm[kPanel];
  __shared__ double pall[kBig];
  __shared__ double wtot[kBig / 32];
  __shared__ int ntot[kBig / 32];
  const int a = threadIdx.x, lane = a & 31, warp = a >> 5;
  const int base = J * kBig;
  const int R = base + a;
  const bool valid = R < n;
  double wv = valid ? w_in[R] : 0.0;
  pall[a] = 0.0;
  const int rr = a >> 3, part = a & 7;  // mat-vec role: row rr of the block, columns 4*part..+3
  // this thread's 32 entries of L[R][c0 .. c0+31] (needed only by rows below the block); the segment of
  // block s+1 is loaded while the mat-vec and the FMAs of block s run
  double lrow[kPanel], lnext[kPanel];
  {
    const bool below0 = valid && R >= base + kPanel;
#pragma unroll
    for (int k = 0; k < kPanel; ++k) lnext[k] = below0 ? (double)L[idx<LAYOUT>(R, base + k, ld)] : 0.0;
  }
  for (int s = 0; s < kBig / kPanel; ++s) {
    const int c0 = base + s * kPanel;
    if (c0 >= n) break;  // uniform
    const double *X = Xp + (int64_t)s * kXBlock + rr * kXStride + part * 4;
    const double x0 = X[0], x1 = X[1], x2 = X[2], x3 = X[3];
    const bool below = valid && R >= c0 + kPanel;
#pragma unroll
    for (int k = 0; k < kPanel; ++k) lrow[k] = lnext[k];
    {
      const int c1 = c0 + kPanel;
      const bool below1 = s + 1 < kBig / kPanel && c1 < n && valid && R >= c1 + kPanel;
#pragma unroll
      for (int k = 0; k < kPanel; ++k) lnext[k] = below1 ? (double)L[idx<LAYOUT>(R, c1 + k, ld)] : 0.0;
    }
    if (warp == s) wsm[lane] = wv;
    __syncthreads();
    double acc = x0 * wsm[part * 4] + x1 * wsm[part * 4 + 1] + x2 * wsm[part * 4 + 2] +
                 x3 * wsm[part * 4 + 3];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (part == 0) pall[s * kPanel + rr] = acc;
    __syncthreads();
    if (below) {
      const double *pp = pall + s * kPanel;
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int k = 0; k < kPanel; k += 2) {
        d0 += lrow[k] * pp[k];
        d1 += lrow[k + 1] * pp[k + 1];
      }
      wv -= d0 + d1;
    }
  }
  __syncthreads();
  const double pk = pall[a];
  if (valid) p_out[a] = pk;
  if (!UPDATE) return;

  // ---- coefficients of column k = R (block scan of p^2 and of the diagonal signs)
  const double lkk = valid ? (double)L[idx<LAYOUT>(R, R, ld)] : 1.0;
  const int neg = lkk < 0.0 ? 1 : 0;
  double incl = pk * pk;
  int nincl = neg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double y = __shfl_up_sync(0xffffffffu, incl, o);
    int m = __shfl_up_sync(0xffffffffu, nincl, o);
    if (lane >= o) { incl += y; nincl += m; }
  }
  if (lane == 31) { wtot[warp] = incl; ntot[warp] = nincl; }
  __syncthreads();
  double off = scal_in[0];
  int noff = 0;
  for (int q = 0; q < warp; ++q) { off += wtot[q]; noff += ntot[q]; }
  const double tk1 = off + incl, tk = tk1 - pk * pk > off ? off + (incl - pk * pk) : off;
  const double S_in = scal_in[1];
  const double Sk = ((noff + nincl - neg) & 1) ? -S_in : S_in;  // prod of signs of columns < k
  const double sgn = neg ? -1.0 : 1.0;
  const double rinv = 1.0 / sqrt(tk * tk1);
  const double ck = sgn * tk * rinv;
  const double sigk = sgn * pk * rinv;
  if (valid) {
    c_out[a] = ck;
    sg_out[a] = sigk;
    dn_out[a] = fabs(lkk) * tk1 * rinv;
    // rotg's z from (a, b) = (L_kk, nu_k), nu_k = S_k L_kk p_k / sqrt(t_k), s_k = S_k sgn_k p_k / sqrt(t_{k+1})
    const double nuk = Sk * lkk * pk / sqrt(tk);
    const double sk = Sk * sgn * pk / sqrt(tk1);
    const double z = rotg_z<double>(lkk, nuk, ck, sk);
    if (v) v[R] = (T)z;
    if (z_out) z_out[a] = z;
  }
  __syncthreads();
  if (a == kBig - 1) {
    scal_out[0] = tk1;
    scal_out[1] = ((noff + nincl) & 1) ? -S_in : S_in;
  }
}

template <typename T, int LAYOUT, bool UPDATE>
__global__ void __launch_bounds__(kBig) large_chain_kernel(const T *L, int n, int64_t ld, T *v,
                                                           LargeWs ws, int J,
                                                           const int32_t *status) {
  if (*status != 0) return;
  const int base = J * kBig;
  chain_body<T, LAYOUT, UPDATE>(L, n, ld, J, ws.X + (int64_t)(base / kPanel) * kXBlock, ws.w, ws.p + base,
                                ws.c + base, ws.sg + base, ws.dn + base, v, (double *)nullptr, ws.scal,
                                ws.scal);
}


// ------------------------------------------------------------- trailing rows
// Rows R >= J*256, columns [J*256, min(n, J*256+256)) and below/at the diagonal:
//   UPDATE:  L_Rk <- c_k L_Rk + sigma_k w ; w <- w - L_Rk(old) p_k ; L_RR <- dn_R
//   !UPDATE: w <- w - L_Rk p_k            (read-only pass of the downdate's trsv)
// Thread = row.  Row-major factors go through a swizzled shared tile (coalesced global access);
// column-major factors are read directly (lanes = consecutive rows).
template <typename T, int LAYOUT, bool UPDATE>
__global__ void __launch_bounds__(128) large_trail_kernel(T *L, int n, int64_t ld, LargeWs ws,
                                                          int J, const int32_t *status) {
  constexpr int NT = 128;
  __shared__ double cs[3][kPanel];
  __shared__ __align__(16) T tile[LAYOUT == CHUB_ROW_MAJOR ? NT * kPanel : 1];
  if (*status != 0) return;
  const int tid = threadIdx.x;
  const int base = J * kBig;
  const int r0 = base + (UPDATE ? 0 : kBig) + blockIdx.x * NT;  // trsv pass: rows below the block only
  if (r0 >= n) return;
  const int rows = min(NT, n - r0);
  const int R = r0 + tid;
  const bool valid = tid < rows;
  double w = valid ? ws.w[R] : 0.0;
  const int cend = min(n, base + kBig);
  for (int c0 = base; c0 < cend; c0 += kPanel) {
    const int bw = min(kPanel, n - c0);
    if (c0 > r0 + rows - 1) break;  // the whole tile is above the diagonal (uniform)
    __syncthreads();
    if (tid < kPanel) {
      cs[0][tid] = tid < bw ? ws.p[c0 + tid] : 0.0;
      if (UPDATE) {
        cs[1][tid] = tid < bw ? ws.c[c0 + tid] : 1.0;
        cs[2][tid] = tid < bw ? ws.sg[c0 + tid] : 0.0;
      }
    }
    if (LAYOUT == CHUB_ROW_MAJOR) {
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
        tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
      }
    } else {
      __syncthreads();
      if (valid) {
        T *q = L + R + (int64_t)c0 * ld;
#pragma unroll 8
        for (int k = 0; k < kPanel; ++k) {
          if (k < bw && c0 + k < R) {
            const double lo = (double)q[(int64_t)k * ld];
            if (UPDATE) q[(int64_t)k * ld] = (T)(cs[1][k] * lo + cs[2][k] * w);
            w -= lo * cs[0][k];
          }
        }
        if (UPDATE && R >= c0 && R < c0 + bw) q[(int64_t)(R - c0) * ld] = (T)ws.dn[R];
      }
    }
  }
  if (valid && R >= cend) ws.w[R] = w;
}

// ------------------------------------------------------------ triangle transpose
// Lower triangle of a column-major matrix <-> row-major scratch, 32x32 tiles through shared memory
// (coalesced on both sides).  TO_ROW: dst[i*ldd + j] = src[i + j*lds] for j <= i;  otherwise the
// inverse (dst[i + j*ldd] = src[i*lds + j]).  Entries above the diagonal are never written.
template <typename T, bool TO_ROW>
__global__ void __launch_bounds__(256) tri_transpose_kernel(T *dst, int64_t ldd, const T *src, int64_t lds, int n) {
  __shared__ T tile[32][33];
  // linear tile index -> (bi >= bj)
  const int t = blockIdx.x;
  int bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((int64_t)bi * (bi + 1) / 2 > t) --bi;
  while ((int64_t)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = t - (int)((int64_t)bi * (bi + 1) / 2);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int i0 = bi * 32, j0 = bj * 32;
  if (TO_ROW) {
    // read column-major: consecutive threads = consecutive rows i
    for (int c = ty; c < 32; c += 8) {
      const int i = i0 + tx, j = j0 + c;
      if (i < n && j < n) tile[c][tx] = src[i + (int64_t)j * lds];
    }
    __syncthreads();
    // write row-major: consecutive threads = consecutive columns j
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, j = j0 + tx;
      if (i < n && j <= i) dst[(int64_t)i * ldd + j] = tile[tx][r];
    }
  } else {
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, j = j0 + tx;
      if (i < n && j < n) tile[r][tx] = src[(int64_t)i * lds + j];
    }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
      const int i = i0 + tx, j = j0 + c;
      if (i < n && j <= i) dst[i + (int64_t)j * ldd] = tile[tx][c];
    }
  }
}

// ------------------------------------------------------- downdate coefficients
// One CTA of 1024 threads: rho^2 = 1 - p.p (.pyx:253-259), suffix sums, (c_k, s_k), sign of the
// input diagonal, v <- rotg's z_k (or v <- p when the result is not positive definite).
// Thread t owns the contiguous chunk [b0, b1) of columns; the chunk sums are combined by a block-wide
// suffix scan (warp shuffles + one shared-memory hop) instead of every thread summing all 1024 partials.
// The same coefficients for the persistent-kernel paths, which keep a copy of the input diagonal (ws.dg): every
// access coalesced, 1024 elements per block step (suffix scan over the block: warp shuffles + one exchange of the
// 32 warp totals), the square roots / divisions of different elements independent of each other.  (The kernel
// below walks 16 consecutive elements per thread with strided loads and a serial sqrt/div chain: 88 us at
// n = 16384, a tenth of the whole downdate -- profiles/r03n.)
template <typename T>
__global__ void __launch_bounds__(1024) dd_coef_par_kernel(int n, T *v, LargeWs ws, int32_t *status) {
  constexpr int NT = 1024;
  __shared__ double wsum[NT / 32];
  __shared__ double wabove[NT / 32];
  __shared__ double total_s;
  if (*status != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nch = (n + NT - 1) / NT;
  // pass 1: p . p
  double acc = 0.0;
  for (int i = tid; i < n; i += NT) { const double x = ws.p[i]; acc = fma(x, x, acc); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    double t = wsum[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) total_s = t;
  }
  __syncthreads();
  const double p_dot_p = total_s;
  const T rho_sq_t = T(1) - T(p_dot_p);
  if (tid == 0) { ws.scal[2] = p_dot_p; ws.scal[3] = (double)rho_sq_t; }
  if (rho_sq_t <= T(0)) {
    for (int i = tid; i < n; i += NT) v[i] = (T)ws.p[i];
    if (tid == 0) *status = CHUB_STATUS_NOT_PD;
    return;
  }
  const double rho_sq = (double)rho_sq_t;
  // pass 2: blocks of 1024 elements from the end; carry = sum of p_j^2 over all later blocks
  double carry = 0.0;
  int k = (nch - 1) * NT + tid;
  double pk = k < n ? ws.p[k] : 0.0, dgk = k < n ? ws.dg[k] : 1.0;
  for (int c = nch - 1; c >= 0; --c) {
    const int kn = k - NT;
    const double pn = (c > 0) ? ws.p[kn] : 0.0, dgn = (c > 0) ? ws.dg[kn] : 1.0;  // next block, in flight
    const double x = pk * pk;
    double suf = x;  // inclusive suffix sum within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += y;
    }
    __syncthreads();  // (wsum / wabove of the previous block have been read)
    if (lane == 0) wsum[warp] = suf;
    __syncthreads();
    if (warp == 0) {
      const double t = wsum[lane];
      double s2 = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_down_sync(0xffffffffu, s2, o);
        if (lane + o < 32) s2 += y;
      }
      wabove[lane] = s2 - t;  // warps strictly above
      if (lane == 0) total_s = s2;
    }
    __syncthreads();
    const double after = carry + wabove[warp] + (suf - x);  // sum_{j > k} p_j^2
    if (k < n) {
      const double qk1 = sqrt(rho_sq + after);
      const double qk = sqrt(rho_sq + (after + x));
      const double cc = qk1 / qk, ss = pk / qk;
      const double sgn = dgk < 0.0 ? -1.0 : 1.0;
      ws.c[k] = cc;
      ws.sg[k] = ss;
      ws.dn[k] = sgn;
      // the same, packed per column for the sweeps that take them in one 32-byte piece: (c, s, sgn c, sgn s).
      // (ws.X is free by now: the forward substitution that used the inverse diagonal blocks is over.)
      double2 *pk4 = reinterpret_cast<double2 *>(ws.X) + 2 * (size_t)k;
      pk4[0] = make_double2(cc, ss);
      pk4[1] = make_double2(sgn * cc, sgn * ss);
      v[k] = rotg_z<T>(T(qk1), T(pk), T(cc), T(ss));
    }
    carry += total_s;
    k = kn; pk = pn; dgk = dgn;
  }
}

// The same on one CTA per 1024 columns: every CTA forms p . p (the rho^2 test: the same reduction order in every
// CTA, so all of them take the same decision) and the sum over the columns behind its own block by itself
// (n doubles out of L2 per CTA), then scans its block -- no CTA waits for another one.  28 -> ~6 us at n = 16384.
template <typename T>
__global__ void __launch_bounds__(1024) dd_coef_par2_kernel(int n, T *v, LargeWs ws, int32_t *status) {
  constexpr int NT = 1024;
  __shared__ double wsum[NT / 32];
  __shared__ double wsum2[NT / 32];
  __shared__ double wabove[NT / 32];
  __shared__ double total_s, behind_s;
  // (NOT_PD can only come from a sibling CTA of this launch: a CTA that starts late still has to take the same
  // branch and write its part of v <- p; every other status means that an earlier kernel gave up)
  const int32_t st0 = *status;
  if (st0 != 0 && st0 != CHUB_STATUS_NOT_PD) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (int)blockIdx.x;
  const int k_end = (c + 1) * NT;  // columns >= k_end are behind my block
  double acc = 0.0, acc2 = 0.0;
  for (int i = tid; i < n; i += NT) {
    const double x = ws.p[i];
    acc = fma(x, x, acc);
    if (i >= k_end) acc2 = fma(x, x, acc2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_down_sync(0xffffffffu, acc, o);
    acc2 += __shfl_down_sync(0xffffffffu, acc2, o);
  }
  if (lane == 0) { wsum[warp] = acc; wsum2[warp] = acc2; }
  __syncthreads();
  if (warp == 0) {
    double t = wsum[lane], t2 = wsum2[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t += __shfl_down_sync(0xffffffffu, t, o);
      t2 += __shfl_down_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) { total_s = t; behind_s = t2; }
  }
  __syncthreads();
  const double p_dot_p = total_s;
  const T rho_sq_t = T(1) - T(p_dot_p);
  const int k = c * NT + tid;
  if (c == 0 && tid == 0) { ws.scal[2] = p_dot_p; ws.scal[3] = (double)rho_sq_t; }
  if (rho_sq_t <= T(0)) {  // not positive definite: v <- p, L stays as it is (the sweep sees the status)
    if (k < n) v[k] = (T)ws.p[k];
    if (tid == 0) *status = CHUB_STATUS_NOT_PD;  // (every CTA writes the same value)
    return;
  }
  const double rho_sq = (double)rho_sq_t;
  const double pk = k < n ? ws.p[k] : 0.0, dgk = k < n ? ws.dg[k] : 1.0;
  const double x = pk * pk;
  double suf = x;  // inclusive suffix sum within the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += y;
  }
  __syncthreads();
  if (lane == 0) wsum[warp] = suf;
  __syncthreads();
  if (warp == 0) {
    const double t = wsum[lane];
    double s2 = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_down_sync(0xffffffffu, s2, o);
      if (lane + o < 32) s2 += y;
    }
    wabove[lane] = s2 - t;  // warps strictly above
  }
  __syncthreads();
  const double after = behind_s + wabove[warp] + (suf - x);  // sum_{j > k} p_j^2
  if (k < n) {
    const double qk1 = sqrt(rho_sq + after);
    const double qk = sqrt(rho_sq + (after + x));
    const double cc = qk1 / qk, ss = pk / qk;
    const double sgn = dgk < 0.0 ? -1.0 : 1.0;
    ws.c[k] = cc;
    ws.sg[k] = ss;
    ws.dn[k] = sgn;
    double2 *pk4 = reinterpret_cast<double2 *>(ws.X) + 2 * (size_t)k;  // packed (c, s, sgn c, sgn s), see above
    pk4[0] = make_double2(cc, ss);
    pk4[1] = make_double2(sgn * cc, sgn * ss);
    v[k] = rotg_z<T>(T(qk1), T(pk), T(cc), T(ss));
  }
}

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(1024) large_dd_coef_kernel(const T *L, int n, int64_t ld, T *v,
                                                             LargeWs ws, int32_t *status) {
  constexpr int NT = 1024;
  __shared__ double wsum[NT / 32];   // per-warp totals
  __shared__ double wabove[NT / 32]; // sum of the totals of the warps above
  __shared__ double total_s;
  if (*status != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = (n + NT - 1) / NT;
  const int b0 = min(n, tid * ch), b1 = min(n, b0 + ch);
  double acc = 0.0;
  for (int i = b0; i < b1; ++i) acc += ws.p[i] * ws.p[i];
  // inclusive suffix sum over the lanes of a warp: suf = sum of acc over lanes >= lane
  double suf = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += y;
  }
  if (lane == 0) wsum[warp] = suf;
  __syncthreads();
  if (warp == 0) {
    double t = wsum[lane];  // NT / 32 == 32 warps
    double s2 = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_down_sync(0xffffffffu, s2, o);
      if (lane + o < 32) s2 += y;
    }
    wabove[lane] = s2 - t;  // warps strictly above
    if (lane == 0) total_s = s2;
  }
  __syncthreads();
  const double p_dot_p = total_s;
  const double above = wabove[warp] + (suf - acc);  // sum over all threads > tid
  const T rho_sq_t = T(1) - T(p_dot_p);
  if (tid == 0) { ws.scal[2] = p_dot_p; ws.scal[3] = (double)rho_sq_t; }
  if (rho_sq_t <= T(0)) {
    for (int i = b0; i < b1; ++i) v[i] = (T)ws.p[i];
    if (tid == 0) *status = CHUB_STATUS_NOT_PD;
    return;
  }
  const double rho_sq = (double)rho_sq_t;
  double run = above;  // sum_{j >= b1} p_j^2
  for (int k = b1 - 1; k >= b0; --k) {
    const double pk = ws.p[k];
    const double qk1 = sqrt(rho_sq + run);
    run += pk * pk;
    const double qk = sqrt(rho_sq + run);
    const double c = qk1 / qk, s = pk / qk;
    ws.c[k] = c;
    ws.sg[k] = s;
    if (L) ws.dn[k] = L[idx<LAYOUT>(k, k, ld)] < T(0) ? -1.0 : 1.0;  // (L == nullptr: ws.dn already holds the signs)
    v[k] = rotg_z<T>(T(qk1), T(pk), T(c), T(s));
  }
}

// ------------------------------------------------------------- downdate sweep
// Row-independent reverse sweep (appendix A.3; .pyx:285-330).  Thread = row; each CTA owns 128
// rows and walks its 32-column tiles from the diagonal to column 0 with x_i in a register.
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(128) large_dd_sweep_kernel(T *L, int n, int64_t ld, LargeWs ws,
                                                             const int32_t *status) {
  constexpr int NT = 128;
  __shared__ double cs[3][kPanel];
  __shared__ __align__(16) T tile[LAYOUT == CHUB_ROW_MAJOR ? NT * kPanel : 1];
  if (*status != 0) return;
  const int tid = threadIdx.x;
  const int nblk = (n + NT - 1) / NT;
  const int r0 = (nblk - 1 - blockIdx.x) * NT;  // longest rows first
  const int rows = min(NT, n - r0);
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
    if (LAYOUT == CHUB_ROW_MAJOR) {
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
    } else {
      __syncthreads();
      if (valid) {
        T *q = L + R + (int64_t)c0 * ld;
#pragma unroll 8
        for (int kk = kPanel - 1; kk >= 0; --kk) {
          if (kk < bw && c0 + kk <= R) {
            const double lo = (double)q[(int64_t)kk * ld], c = cs[0][kk], s = cs[1][kk];
            q[(int64_t)kk * ld] = (T)(cs[2][kk] * (c * lo - s * x));
            x = c * x + s * lo;
          }
        }
      }
    }
  }
}

}  // namespace chub
