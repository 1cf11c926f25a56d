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

__global__ void rc_pack_kernel(LargeWs ws, int J, double *buf, const int32_t *status) {
  const int t = threadIdx.x;  // kBig threads
  const int k = J * kBig + t;
  buf[t] = ws.p[k];
  buf[kBig + t] = ws.c[k];
  buf[2 * kBig + t] = ws.sg[k];
  buf[3 * kBig + t] = ws.dn[k];
  if (t < 7) buf[4 * kBig + t] = ws.scal[t];
  if (t == 7) buf[4 * kBig + 7] = (double)*status;  // the owner's status travels with the panel
}

__global__ void rc_unpack_kernel(LargeWs ws, int J, const double *buf, int32_t *status) {
  const int t = threadIdx.x;
  const int k = J * kBig + t;
  ws.p[k] = buf[t];
  ws.c[k] = buf[kBig + t];
  ws.sg[k] = buf[2 * kBig + t];
  ws.dn[k] = buf[3 * kBig + t];
  if (t < 7) ws.scal[t] = buf[4 * kBig + t];
  if (t == 7 && buf[4 * kBig + 7] != 0.0) atomicMax(status, (int)buf[4 * kBig + 7]);
}

// Trailing update of the local rows whose global index is >= J*256 (same arithmetic as
// large_trail_kernel<T, ROW_MAJOR, true>; one CTA = 128 consecutive local = global rows).
template <typename T>
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
  if (rows <= 0 || g0 + rows <= base) return;  // entirely above the panel
  const int r0 = (int)g0;
  T *L = A + ((int64_t)l0 - g0) * ld;  // so that L[R * ld + c] addresses global row R
  const int R = r0 + tid;
  const bool valid = tid < rows && R >= base;
  double w = valid ? ws.w[R] : 0.0;
  const int cend = min(n, base + kBig);
  for (int c0 = base; c0 < cend; c0 += kPanel) {
    const int bw = min(kPanel, n - c0);
    if (c0 > r0 + rows - 1) break;
    __syncthreads();
    if (tid < kPanel) {
      cs[0][tid] = tid < bw ? ws.p[c0 + tid] : 0.0;
      cs[1][tid] = tid < bw ? ws.c[c0 + tid] : 1.0;
      cs[2][tid] = tid < bw ? ws.sg[c0 + tid] : 0.0;
    }
    tile_load_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
    __syncthreads();
    if (valid) {
#pragma unroll 8
      for (int k = 0; k < kPanel; ++k) {
        if (k < bw && c0 + k < R) {
          const int o = tile_off(tid, k);
          const double lo = (double)tile[o];
          tile[o] = (T)(cs[1][k] * lo + cs[2][k] * w);
          w -= lo * cs[0][k];
        }
      }
      if (R >= c0 && R < c0 + bw) tile[tile_off(tid, R - c0)] = (T)ws.dn[R];
    }
    __syncthreads();
    // rows above the panel inside this CTA (R < base) are loaded but never stored back changed
    tile_store_rm<T, NT>(tile, L, ld, r0, rows, c0, bw);
  }
  if (valid && R >= cend) ws.w[R] = w;
}

}  // namespace chub
