// small.cuh -- one CTA per factor: the literal Seeger algorithm (Givens rotations,
// column by column) for small n and for batches of independent factors.
//
// Restates _seeger_impl_cython.pyx:92-155 (update) and :224-330 (downdate) in the
// row-independent form of SURVEY appendix A.2 / A.3: within a panel of 32 columns
// warp 0 walks the 32 x 32 diagonal block column by column (the sequential part:
// rotg + sign fix), publishing (c_k, s_k) to shared memory; then every row below
// the block applies the 32 rotations to its own entries with its running v_i in a
// register (BLAS rot, x' = c x + s y, y' = c y - s x, _blas_polymorphic.pyx:19-31).
// Row-major factors are staged through an XOR-swizzled shared-memory tile so that
// global traffic is coalesced; column-major factors are accessed directly (lanes =
// consecutive rows).  Only i >= j is ever read into arithmetic or written.
#pragma once

#include "common.cuh"

namespace chub {

template <typename T>
struct SmallArgs {
  T *L;
  T *v;
  int32_t *status;  // [batch]
  int n;
  int64_t ld, stride_L, stride_v;
  int flags;
};

constexpr int kSmallNT = 128;    // threads per CTA (= rows per staged tile)
constexpr int kSmallNTBatched = 32;     // large batches, update: one warp per factor (launch_small, api.cu)
constexpr int kSmallNTBatchedDd = 64;   // large batches, downdate: two warps per factor (measured best)
constexpr int kSmallMaxN = 2048;  // shared-memory bound of the downdate kernel

// L2 prefetch of the 32-column segment [pc, pc + 32) of row pr (row-major), if it exists.
template <typename T>
__device__ __forceinline__ void prefetch_row_l2(const T *L, int64_t ld, int n, int pr, int pc) {
  if (pr < n && pc < n) {
    const T *q = L + (int64_t)pr * ld + pc;
    prefetch_l2(q);
    if (kPanel * sizeof(T) > 128 && pc + 16 < n) prefetch_l2(q + 16);
  }
}

template <typename T, int LAYOUT, int NT>
__device__ __forceinline__ bool small_zero_diag(const T *L, int n, int64_t ld) {
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += NT) bad |= (L[idx<LAYOUT>(i, i, ld)] == T(0));
  return __syncthreads_or(bad) != 0;
}

// ------------------------------------------------------------------ update
template <typename T, int LAYOUT, int NT>
__global__ void __launch_bounds__(NT) small_update_kernel(SmallArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n;
  const int64_t ld = a.ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T *L = a.L + (int64_t)blockIdx.x * a.stride_L;
  T *v = a.v + (int64_t)blockIdx.x * a.stride_v;
  const int n_pad = (n + 31) & ~31;
  T *nu = reinterpret_cast<T *>(smem_raw);  // running v (A.2: nu_i), [n_pad]
  T *cc = nu + n_pad;                       // panel cosines [32]
  T *ss = cc + kPanel;                      // panel sines   [32]
  T *pp = ss + kPanel;                      // block p = L_D^{-1} nu [32]
  T *sgm = pp + kPanel;                     // sigma_k [32]
  T *tile = sgm + kPanel;                   // [NT][32], row-major staging only

  if (a.flags & CHUB_FLAG_CHECK_DIAG) {  // _arg_validation.py:18-22
    if (small_zero_diag<T, LAYOUT, NT>(L, n, ld)) {
      if (tid == 0) a.status[blockIdx.x] = CHUB_STATUS_ZERO_DIAG;
      return;
    }
  }
  for (int i = tid; i < n; i += NT) nu[i] = v[i];
  if (LAYOUT == CHUB_ROW_MAJOR) prefetch_row_l2<T>(L, ld, n, kPanel + tid, 0);  // first trailing tile
  __syncthreads();

  for (int c0 = 0; c0 < n; c0 += kPanel) {
    const int bw = min(kPanel, n - c0);
    if (warp == 0) {
      // Diagonal block: lane = row c0 + lane, its entries L[row][c0 .. row] in registers.
      // The 32 rotations of the block are formed WITHOUT a sqrt / division per column on the
      // sequential chain, through their closed form (restarted per block, t_0 = 1):
      //   p = L_D^{-1} nu (forward substitution: one multiply, one shuffle, one FMA per column),
      //   t_{k+1} = t_k + p_k^2 (warp scan),  c_k = sgn_k sqrt(t_k/t_{k+1}),
      //   s_k = S_k sgn_k p_k / sqrt(t_{k+1}),  L'_kk = |L_kk| sqrt(t_{k+1}/t_k),
      //   L'_ik = c_k L_ik + sgn_k p_k/sqrt(t_k t_{k+1}) w_i^{(k)}   (== c_k L_ik + s_k nu_i^{(k)}),
      // which is rotg + sign fix + rot of .pyx:101-140 for these 32 columns.
      const int row = c0 + lane;
      const bool valid = lane < bw;
      T Lrow[kPanel];
      T dii = T(1);
#pragma unroll
      for (int k = 0; k < kPanel; ++k) {
        Lrow[k] = (valid && k <= lane) ? L[idx<LAYOUT>(row, c0 + k, ld)] : T(0);
        if (k == lane && valid) dii = Lrow[k];
      }
      const T w0 = valid ? nu[row] : T(0);
      const T rinv = T(1) / dii;
      T w = w0, pk_mine = T(0);
#pragma unroll
      for (int k = 0; k < kPanel; ++k) {
        if (k < bw) {  // uniform
          // branch-free: Lrow[k] == 0 for lanes above the diagonal (k > lane), and lane k's own w
          // is dead after its column, so the update runs unconditionally (no divergent branches)
          const T pk = shfl(w * rinv, k);
          pk_mine = (lane == k) ? pk : pk_mine;
          w -= Lrow[k] * pk;
        }
      }
      // coefficients of all columns of the block at once (fp64 arithmetic)
      const double pd = (double)pk_mine;
      double incl = pd * pd;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 0.0;
      const double tk = 1.0 + excl, tk1 = 1.0 + incl;
      const bool neg = dii < T(0);
      const unsigned negmask = __ballot_sync(0xffffffffu, neg);
      const double Sk = (__popc(negmask & ((1u << lane) - 1u)) & 1) ? -1.0 : 1.0;
      const double sgn = neg ? -1.0 : 1.0;
      const double isq = rsqrt(tk), r1 = rsqrt(tk1);
      const double ck = sgn * tk * isq * r1;          // sgn sqrt(t_k / t_{k+1})
      const double sigk = sgn * pd * isq * r1;        // sgn p_k / sqrt(t_k t_{k+1})
      const double sk = Sk * sgn * pd * r1;           // rotation sine seen by the rows below
      const double dnew = fabs((double)dii) * tk1 * isq * r1;
      cc[lane] = (T)ck;
      ss[lane] = (T)sk;
      pp[lane] = pk_mine;
      sgm[lane] = (T)sigk;
      __syncwarp();
      // second pass (lane-parallel): rebuild w_i^{(k)} and rotate the block's own entries
      T w2 = w0;
#pragma unroll
      for (int k = 0; k < kPanel; ++k) {
        if (k < bw) {  // uniform; entries with k >= lane get garbage that is never stored
          const T x = Lrow[k];
          Lrow[k] = cc[k] * x + sgm[k] * w2;
          w2 -= x * pp[k];
        }
      }
      if (valid) {
#pragma unroll
        for (int k = 0; k < kPanel; ++k) {
          if (k < lane) L[idx<LAYOUT>(row, c0 + k, ld)] = Lrow[k];
          else if (k == lane) L[idx<LAYOUT>(row, c0 + k, ld)] = (T)dnew;
        }
        // what rotg leaves in v[k]: z from (a, b) = (L_kk, nu_k)
        const double nuk = Sk * (double)dii * pd * isq;
        v[row] = (T)rotg_z<double>((double)dii, nuk, ck, sk);
      }
    }
    __syncthreads();

    // Rows below the diagonal block: thread = row, serial over the panel's columns.
    const int rbeg = c0 + kPanel;
    if (LAYOUT == CHUB_ROW_MAJOR) {
      for (int r0 = rbeg; r0 < n; r0 += NT) {
        const int rows = min(NT, n - r0);
        tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, kPanel);
        {
          // The tiles of a factor are visited one after the other with a DRAM round trip each: pull the
          // next tile (same panel, or the next panel's diagonal block + first trailing rows) into L2
          // while this one is processed.  thread = row of that tile.
          const bool last = r0 + NT >= n;
          const int pc = last ? c0 + kPanel : c0;
          for (int pr = (last ? rbeg : r0 + NT) + tid, rep = 0; rep < (last && NT < 64 ? 2 : 1); ++rep, pr += NT)
            prefetch_row_l2<T>(L, ld, n, pr, pc);
        }
        __syncthreads();
        if (tid < rows) {
          T w = nu[r0 + tid];
#pragma unroll 8
          for (int k = 0; k < kPanel; ++k) {
            const int o = tile_off(tid, k);
            T x = tile[o], c = cc[k], s = ss[k];
            tile[o] = c * x + s * w;
            w = c * w - s * x;
          }
          nu[r0 + tid] = w;
        }
        __syncthreads();
        tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, kPanel);
        __syncthreads();
      }
    } else {
      for (int i = rbeg + tid; i < n; i += NT) {
        T w = nu[i];
        T *p = L + i + (int64_t)c0 * ld;
#pragma unroll 8
        for (int k = 0; k < kPanel; ++k) {
          T x = p[(int64_t)k * ld], c = cc[k], s = ss[k];
          p[(int64_t)k * ld] = c * x + s * w;
          w = c * w - s * x;
        }
        nu[i] = w;
      }
    }
    __syncthreads();
  }
}

template <typename T>
inline size_t small_update_smem(int n, int layout, int nt = kSmallNT) {
  size_t n_pad = (n + 31) & ~31;
  return (n_pad + 4 * kPanel + (layout == CHUB_ROW_MAJOR ? (size_t)nt * kPanel : 0)) * sizeof(T);
}

// ---------------------------------------------------------------- downdate
template <typename T, int LAYOUT, int NT>
__global__ void __launch_bounds__(NT) small_downdate_kernel(SmallArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n;
  const int64_t ld = a.ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T *L = a.L + (int64_t)blockIdx.x * a.stride_L;
  T *v = a.v + (int64_t)blockIdx.x * a.stride_v;
  const int n_pad = (n + 31) & ~31;
  double *suf = reinterpret_cast<double *>(smem_raw);  // [n_pad + 2] suffix sums of p^2
  double *red = suf + n_pad + 2;                         // [NT] reduction scratch
  T *p = reinterpret_cast<T *>(red + NT);                // [n_pad]  v -> p = L^{-1} v
  T *cc = p + n_pad;                                     // [n_pad]
  T *ss = cc + n_pad;                                    // [n_pad]
  T *sg = ss + n_pad;                                    // [n_pad] sign of the input diagonal
  T *tile = sg + n_pad;                                  // [NT][32], row-major staging only
  T *xs = reinterpret_cast<T *>(suf);                    // aliases suf after the coefficients exist

  if (a.flags & CHUB_FLAG_CHECK_DIAG) {
    if (small_zero_diag<T, LAYOUT, NT>(L, n, ld)) {
      if (tid == 0) a.status[blockIdx.x] = CHUB_STATUS_ZERO_DIAG;
      return;
    }
  }
  for (int i = tid; i < n; i += NT) p[i] = v[i];
  __syncthreads();

  // 1. p = L^{-1} v: blocked forward substitution (trsv, .pyx:239-248)
  for (int c0 = 0; c0 < n; c0 += kPanel) {
    const int bw = min(kPanel, n - c0);
    if (warp == 0) {
      const int row = c0 + lane;
      const bool valid = lane < bw;
      T Lrow[kPanel];
#pragma unroll
      for (int k = 0; k < kPanel; ++k)
        Lrow[k] = (valid && k <= lane) ? L[idx<LAYOUT>(row, c0 + k, ld)] : T(0);
      T w = valid ? p[row] : T(0);
      T dii = T(1);
#pragma unroll
      for (int k = 0; k < kPanel; ++k)
        if (k == lane && valid) dii = Lrow[k];
      const T rinv = T(1) / dii;
      T pmine = T(0);
#pragma unroll
      for (int k = 0; k < kPanel; ++k) {
        if (k < bw) {  // branch-free (see the update kernel)
          const T pk = shfl(w * rinv, k);
          pmine = (lane == k) ? pk : pmine;
          w -= Lrow[k] * pk;
        }
      }
      if (valid) p[row] = pmine;
    }
    __syncthreads();
    const int rbeg = c0 + kPanel;
    if (LAYOUT == CHUB_ROW_MAJOR) {
      for (int r0 = rbeg; r0 < n; r0 += NT) {
        const int rows = min(NT, n - r0);
        tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, kPanel);
        {  // next tile -> L2 (same panel, else the next panel's diagonal block and first trailing rows)
          const bool last = r0 + NT >= n;
          prefetch_row_l2<T>(L, ld, n, (last ? rbeg : r0 + NT) + tid, last ? c0 + kPanel : c0);
        }
        __syncthreads();
        if (tid < rows) {
          T w = p[r0 + tid];
#pragma unroll 8
          for (int k = 0; k < kPanel; ++k) w -= tile[tile_off(tid, k)] * p[c0 + k];
          p[r0 + tid] = w;
        }
        __syncthreads();
      }
    } else {
      for (int i = rbeg + tid; i < n; i += NT) {
        T w = p[i];
        const T *q = L + i + (int64_t)c0 * ld;
#pragma unroll 8
        for (int k = 0; k < kPanel; ++k) w -= q[(int64_t)k * ld] * p[c0 + k];
        p[i] = w;
      }
    }
    __syncthreads();
  }

  // 2. rho^2 = 1 - p.p  (.pyx:253-259); the chunked sums double as the scan's block sums
  const int ch = (n + NT - 1) / NT;  // contiguous chunk per thread
  {
    double acc = 0.0;
    const int b0 = tid * ch, b1 = min(n, b0 + ch);
    for (int i = b0; i < b1; ++i) acc += (double)p[i] * (double)p[i];
    red[tid] = acc;
  }
  __syncthreads();
  double p_dot_p = 0.0;
  for (int t = 0; t < NT; ++t) p_dot_p += red[t];
  const T rho_sq_t = T(1) - T(p_dot_p);
  if (rho_sq_t <= T(0)) {  // not positive definite: L untouched, v <- p
    for (int i = tid; i < n; i += NT) v[i] = p[i];
    if (tid == 0) a.status[blockIdx.x] = CHUB_STATUS_NOT_PD;
    return;
  }
  const double rho_sq = (double)rho_sq_t;

  // 3. q_k^2 = rho^2 + sum_{j >= k} p_j^2 (A.3, a sum of non-negatives);
  //    c_k = q_{k+1} / q_k, s_k = p_k / q_k  == rotg(q_np1, p[k]) + sign fix (.pyx:288-298)
  {
    double above = 0.0;  // sum of the chunks to the right of mine
    for (int t = tid + 1; t < NT; ++t) above += red[t];
    const int b0 = tid * ch, b1 = min(n, b0 + ch);
    if (b0 < n) {
      double run = above;
      for (int i = b1 - 1; i >= b0; --i) {
        run += (double)p[i] * (double)p[i];
        suf[i] = run;
      }
    }
    if (tid == 0) suf[n] = 0.0;
  }
  __syncthreads();
  for (int k = tid; k < n; k += NT) {
    const double qk = sqrt(rho_sq + suf[k]), qk1 = sqrt(rho_sq + suf[k + 1]);
    const T c = T(qk1 / qk), s = T((double)p[k] / qk);
    cc[k] = c;
    ss[k] = s;
    sg[k] = L[idx<LAYOUT>(k, k, ld)] < T(0) ? T(-1) : T(1);  // .pyx:319-326
    v[k] = rotg_z(T(qk1), p[k], c, s);                        // what rotg leaves in p[k]
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) xs[i] = T(0);  // aux column of [L | 0]  (.pyx:271-276)
  __syncthreads();

  // 4. sweep k = N-1 .. 0 (.pyx:285-330), row-independent (A.3):
  //    (L_ik, x_i) <- (c_k L_ik - s_k x_i, c_k x_i + s_k L_ik), x_i = 0 at k = i.
  for (int c0 = ((n - 1) / kPanel) * kPanel; c0 >= 0; c0 -= kPanel) {
    const int bw = min(kPanel, n - c0);
    if (LAYOUT == CHUB_ROW_MAJOR) {
      for (int r0 = c0; r0 < n; r0 += NT) {
        const int rows = min(NT, n - r0);
        tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
        {  // next tile -> L2 (same panel, else the top of the panel to the left)
          const bool last = r0 + NT >= n;
          if (!last) prefetch_row_l2<T>(L, ld, n, r0 + NT + tid, c0);
          else if (c0 >= kPanel) prefetch_row_l2<T>(L, ld, n, c0 - kPanel + tid, c0 - kPanel);
        }
        __syncthreads();
        if (tid < rows) {
          const int i = r0 + tid;
          T x = xs[i];
#pragma unroll 8
          for (int kk = kPanel - 1; kk >= 0; --kk) {
            const int k = c0 + kk;
            if (kk < bw && k <= i) {
              const int o = tile_off(tid, kk);
              T lo = tile[o], c = cc[k], s = ss[k];
              tile[o] = sg[k] * (c * lo - s * x);
              x = c * x + s * lo;
            }
          }
          xs[i] = x;
        }
        __syncthreads();
        tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
        __syncthreads();
      }
    } else {
      for (int i = c0 + tid; i < n; i += NT) {
        T x = xs[i];
        T *q = L + i + (int64_t)c0 * ld;
#pragma unroll 8
        for (int kk = kPanel - 1; kk >= 0; --kk) {
          const int k = c0 + kk;
          if (kk < bw && k <= i) {
            T lo = q[(int64_t)kk * ld], c = cc[k], s = ss[k];
            q[(int64_t)kk * ld] = sg[k] * (c * lo - s * x);
            x = c * x + s * lo;
          }
        }
        xs[i] = x;
      }
      __syncthreads();  // row -> thread ownership shifts with c0; xs[] is handed over
    }
  }
}

template <typename T>
inline size_t small_downdate_smem(int n, int layout, int nt = kSmallNT) {
  size_t n_pad = (n + 31) & ~31;
  return (n_pad + 2 + nt) * sizeof(double) +
         (4 * n_pad + (layout == CHUB_ROW_MAJOR ? (size_t)nt * kPanel : 0)) * sizeof(T);
}

}  // namespace chub
