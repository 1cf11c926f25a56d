// host.cu -- host-buffer entry points (chub_*_host_*): what a NumPy caller of the reference
// holds (_seeger.py:37-44) is staged to the device, updated there and copied back.
//
// Only the lower triangle travels: the matrix is cut into row blocks (row-major) or column
// blocks (column-major) and each block is copied as a 2-D rectangle that ends at (starts at)
// the diagonal, so PCIe carries ~n^2/2 elements each way instead of n^2.  The strict upper
// triangle of the host array is never modified.

#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace {

struct HostCache {
  void *L = nullptr;
  size_t L_bytes = 0;
  void *v = nullptr;
  size_t v_bytes = 0;
  int32_t *status = nullptr;
  cudaStream_t st = nullptr;
};
std::mutex h_mu;
HostCache h_cache[64];
thread_local char h_err[256] = "";

#define HCK(call)                                                             \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      snprintf(h_err, sizeof(h_err), "%s: %s", #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                         \
    }                                                                         \
  } while (0)

int ensure(HostCache &c, size_t Lb, size_t vb) {
  if (!c.st) HCK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
  if (!c.status) HCK(cudaMalloc((void **)&c.status, sizeof(int32_t)));
  if (c.L_bytes < Lb) {
    if (c.L) HCK(cudaFree(c.L));
    c.L = nullptr; c.L_bytes = 0;
    HCK(cudaMalloc(&c.L, Lb));
    c.L_bytes = Lb;
  }
  if (c.v_bytes < vb) {
    if (c.v) HCK(cudaFree(c.v));
    c.v = nullptr; c.v_bytes = 0;
    HCK(cudaMalloc(&c.v, vb));
    c.v_bytes = vb;
  }
  return 0;
}

// Copy the lower triangle between host and device in blocks of `blk` rows / columns.
int copy_tri(char *dst, const char *src, int64_t n, int64_t ld_dst, int64_t ld_src, int layout,
             size_t es, cudaMemcpyKind kind, cudaStream_t st) {
  const int64_t blk = 1024;
  for (int64_t b0 = 0; b0 < n; b0 += blk) {
    const int64_t b1 = b0 + blk < n ? b0 + blk : n;
    if (layout == CHUB_ROW_MAJOR) {
      // rows [b0, b1), columns [0, b1)
      HCK(cudaMemcpy2DAsync(dst + (size_t)b0 * ld_dst * es, (size_t)ld_dst * es,
                            src + (size_t)b0 * ld_src * es, (size_t)ld_src * es, (size_t)b1 * es,
                            (size_t)(b1 - b0), kind, st));
    } else {
      // columns [b0, b1), rows [b0, n)
      HCK(cudaMemcpy2DAsync(dst + ((size_t)b0 * ld_dst + b0) * es, (size_t)ld_dst * es,
                            src + ((size_t)b0 * ld_src + b0) * es, (size_t)ld_src * es,
                            (size_t)(n - b0) * es, (size_t)(b1 - b0), kind, st));
    }
  }
  return 0;
}

typedef int (*dev_fn)(void *, int64_t, int64_t, int, void *, int, void *, size_t, int32_t *, void *);

int run_host(dev_fn fn, size_t es, void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
             int32_t *status) {
  if (!status || n < 0 || ld < n) return (int)cudaErrorInvalidValue;
  *status = 0;
  if (n == 0) return 0;
  int dev = 0;
  HCK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(h_mu);
  HostCache &c = h_cache[dev & 63];
  int rc = ensure(c, (size_t)n * n * es, (size_t)n * es);
  if (rc) return rc;
  // device copy is dense (ld = n)
  rc = copy_tri((char *)c.L, (const char *)L, n, n, ld, layout, es, cudaMemcpyHostToDevice, c.st);
  if (rc) return rc;
  HCK(cudaMemcpyAsync(c.v, v, (size_t)n * es, cudaMemcpyHostToDevice, c.st));
  rc = fn(c.L, n, n, layout, c.v, flags, nullptr, 0, c.status, c.st);
  if (rc) {
    snprintf(h_err, sizeof(h_err), "%s", chub_last_error());
    return rc;
  }
  HCK(cudaMemcpyAsync(status, c.status, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));
  HCK(cudaStreamSynchronize(c.st));
  if (*status == CHUB_STATUS_ZERO_DIAG) return 0;  // L and v untouched
  HCK(cudaMemcpyAsync(v, c.v, (size_t)n * es, cudaMemcpyDeviceToHost, c.st));
  if (*status == CHUB_STATUS_OK) {
    rc = copy_tri((char *)L, (const char *)c.L, n, ld, n, layout, es, cudaMemcpyDeviceToHost, c.st);
    if (rc) return rc;
  }
  HCK(cudaStreamSynchronize(c.st));
  return 0;
}

}  // namespace

extern "C" {

int chub_update_host_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_update_f64, 8, L, n, ld, layout, v, flags, status);
}
int chub_update_host_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_update_f32, 4, L, n, ld, layout, v, flags, status);
}
int chub_downdate_host_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_downdate_f64, 8, L, n, ld, layout, v, flags, status);
}
int chub_downdate_host_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_downdate_f32, 4, L, n, ld, layout, v, flags, status);
}

int chub_host_release(void) {
  int dev = 0;
  HCK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(h_mu);
  HostCache &c = h_cache[dev & 63];
  if (c.L) cudaFree(c.L);
  if (c.v) cudaFree(c.v);
  c.L = c.v = nullptr;
  c.L_bytes = c.v_bytes = 0;
  return 0;
}

}  // extern "C"
