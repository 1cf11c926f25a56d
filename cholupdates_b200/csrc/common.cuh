// common.cuh -- shared device helpers for the sm_100a Seeger kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cholupdates_b200.h"

namespace chub {

constexpr int kWarp = 32;
constexpr int kPanel = 32;  // columns per panel / rows per diagonal block

// Element (i, j) of an n x n matrix with leading dimension ld.
template <int LAYOUT>
__device__ __forceinline__ int64_t idx(int64_t i, int64_t j, int64_t ld) {
  return LAYOUT == CHUB_ROW_MAJOR ? i * ld + j : i + j * ld;
}

// Shared-memory tile of [rows][32] elements, XOR-swizzled so that both access
// shapes are bank-conflict free: "warp = one row, lanes = 32 columns" (global
// <-> shared staging of row-major data) and "lanes = 32 consecutive rows, same
// column" (thread-per-row compute).
__device__ __forceinline__ int tile_off(int r, int c) { return r * kPanel + (c ^ (r & 31)); }

// Element-sized cp.async (4 or 8 bytes): global -> shared without a register round trip, so that
// all of a thread's loads are in flight at once.
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}

// global (row-major) -> swizzled shared tile of [rows][32]; complete after the caller's
// __syncthreads() (the wait for this thread's copies is inside).  Warp w moves rows w, w + NT/32, ...
// (lane = column: one coalesced 32-element row segment per instruction, pointer stepped by whole rows).
template <typename T, int NT>
__device__ __forceinline__ void tile_load_rm(T *tile, const T *L, int64_t ld, int r0, int rows,
                                             int c0, int bw) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane < bw) {
    const T *g = L + (int64_t)(r0 + w) * ld + c0 + lane;
    const int64_t step = (int64_t)NW * ld;
#pragma unroll 8
    for (int rr = w; rr < rows; rr += NW, g += step) cp_async_elem(tile + tile_off(rr, lane), g);
  }
  cp_async_wait_all();
}
// swizzled shared tile -> global (row-major); never writes above the diagonal
template <typename T, int NT>
__device__ __forceinline__ void tile_store_rm(const T *tile, T *L, int64_t ld, int r0, int rows,
                                              int c0, int bw) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane < bw) {
    T *g = L + (int64_t)(r0 + w) * ld + c0 + lane;
    const int64_t step = (int64_t)NW * ld;
    const int dlim = c0 + lane - r0;  // store row rr only if rr >= dlim  (column <= row)
#pragma unroll 8
    for (int rr = w; rr < rows; rr += NW, g += step)
      if (rr >= dlim) *g = tile[tile_off(rr, lane)];
  }
}

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x);
template <> __device__ __forceinline__ double t_abs<double>(double x) { return fabs(x); }
template <> __device__ __forceinline__ float t_abs<float>(float x) { return fabsf(x); }

template <typename T> struct Lim;
template <> struct Lim<double> {
  static __device__ __forceinline__ double big() { return 1e150; }
  static __device__ __forceinline__ double tiny() { return 1e-150; }
};
template <> struct Lim<float> {
  static __device__ __forceinline__ float big() { return 1e18f; }
  static __device__ __forceinline__ float tiny() { return 1e-18f; }
};

// Net effect of BLAS xROTG followed by the reference's sign fix
// (_seeger_impl_cython.pyx:101-113):
//   r = +hypot(a, b),  c = a / r,  s = b / r      (c = 1, s = 0, r = 0 if a = b = 0).
template <typename T>
__device__ __forceinline__ void givens(T a, T b, T &c, T &s, T &r) {
  T aa = t_abs(a), ab = t_abs(b);
  T m = aa > ab ? aa : ab;
  if (m == T(0)) {
    c = T(1); s = T(0); r = T(0);
    return;
  }
  if (m < Lim<T>::big() && m > Lim<T>::tiny()) {
    r = t_sqrt(a * a + b * b);
    T inv = T(1) / r;
    c = a * inv;
    s = b * inv;
  } else {  // scaled form, as reference xROTG does, to avoid overflow / underflow
    T as = a / m, bs = b / m;
    T rs = t_sqrt(as * as + bs * bs);
    r = m * rs;
    c = as / rs;
    s = bs / rs;
  }
}

// BLAS xROTG's reconstruction parameter z (what the Cython path leaves in v[k]),
// from the inputs (a, b) and the post-sign-fix rotation (c, s) with r > 0:
//   roe = |a| > |b| ? a : b;  r0 = sign(roe) r;  c0 = a / r0;  s0 = b / r0;
//   z = |a| > |b| ? s0 : (c0 != 0 ? 1 / c0 : 1)             (z = 0 if a = b = 0).
template <typename T>
__device__ __forceinline__ T rotg_z(T a, T b, T c, T s) {
  T aa = t_abs(a), ab = t_abs(b);
  if (aa == T(0) && ab == T(0)) return T(0);
  T roe = aa > ab ? a : b;
  T sg = roe < T(0) ? T(-1) : T(1);
  if (aa > ab) return sg * s;
  T c0 = sg * c;
  return c0 != T(0) ? T(1) / c0 : T(1);
}

// CTA-scope acquire / release on shared-memory progress counters (cheaper than a MEMBAR.CTA fence)
__device__ __forceinline__ int ld_acquire_cta_smem(const volatile int *p) {
  int x;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];\n" : "=r"(x) : "r"((unsigned)__cvta_generic_to_shared(const_cast<const int *>(p))) : "memory");
  return x;
}
__device__ __forceinline__ void st_release_cta_smem(volatile int *p, int x) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(const_cast<int *>(p))), "r"(x) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double shfl(double x, int src) { return __shfl_sync(0xffffffffu, x, src); }
__device__ __forceinline__ float shfl(float x, int src) { return __shfl_sync(0xffffffffu, x, src); }

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

}  // namespace chub
