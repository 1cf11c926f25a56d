// host.cu -- host-buffer entry points (chub_*_host_*): what a NumPy caller of the reference
// holds (_seeger.py:37-44) is staged to the device, updated there and copied back.
//
// Only the lower triangle travels: the matrix is cut into row blocks (row-major) or column
// blocks (column-major) and each block is copied as a 2-D rectangle that ends at (starts at)
// the diagonal, so PCIe carries ~n^2/2 elements each way instead of n^2.  The strict upper
// triangle of the host array is never modified.

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "large.cuh"

namespace {

struct HostCache {
  void *L = nullptr;
  size_t L_bytes = 0;
  void *v = nullptr;
  size_t v_bytes = 0;
  void *ws = nullptr;
  size_t ws_bytes = 0;
  int32_t *status = nullptr;
  cudaStream_t st = nullptr;      // compute
  cudaStream_t st_in = nullptr;   // host -> device copies
  cudaStream_t st_out = nullptr;  // device -> host copies
  std::vector<cudaEvent_t> ev_in, ev_done;
  // pinned bounce buffers for PAGEABLE caller memory (two per direction), see run_host_update_streamed
  void *bin[2] = {nullptr, nullptr}, *bout[2] = {nullptr, nullptr};
  size_t b_bytes = 0;
  cudaEvent_t ev_bin[2] = {nullptr, nullptr}, ev_bout[2] = {nullptr, nullptr};
};
std::mutex h_mu[64];  // one per device: host calls on different devices of a process do not serialise
HostCache h_cache[64];
thread_local char h_err[256] = "";
}  // namespace
extern "C" void chub_internal_set_error(const char *msg);  // api.cu: what chub_last_error() returns
namespace {

// A failing call reports through chub_last_error() and leaves no copy into the caller's arrays in flight.
#define HCK(call)                                                             \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      snprintf(h_err, sizeof(h_err), "%s: %s", #call, cudaGetErrorString(e_)); \
      chub_internal_set_error(h_err);                                         \
      cudaDeviceSynchronize();                                                \
      return (int)e_;                                                         \
    }                                                                         \
  } while (0)

int ensure(HostCache &c, size_t Lb, size_t vb) {
  if (!c.st) HCK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
  if (!c.st_in) HCK(cudaStreamCreateWithFlags(&c.st_in, cudaStreamNonBlocking));
  if (!c.st_out) HCK(cudaStreamCreateWithFlags(&c.st_out, cudaStreamNonBlocking));
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

// ---- pageable caller memory.  A cudaMemcpy2DAsync straight on pageable memory is staged by the driver,
// synchronously and at ~10 GB/s (n = 16384: 183 ms per update against 24 ms from pinned memory).  What a
// drop-in caller passes IS pageable (an ordinary numpy.ndarray), so the streamed update bounces through
// two pinned buffers per direction: host threads pack panel P+1 while the DMA engine moves panel P, and an
// unpacking thread scatters panel P-1 back into the caller's array.
bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int host_threads() {
  static int n = 0;
  if (!n) {
    if (const char *e = getenv("CHUB_HOST_THREADS")) n = atoi(e);
    if (n <= 0) n = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
  }
  return n;
}

// rows [0, height) of `width` bytes: strided (pitch) <-> packed, split over host threads
void copy_rows_parallel(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t height) {
  const int nt = (int)std::min<size_t>(host_threads(), std::max<size_t>(1, height * width / (1 << 20)));
  auto work = [=](size_t r0, size_t r1) {
    for (size_t r = r0; r < r1; ++r) memcpy(dst + r * dpitch, src + r * spitch, width);
  };
  if (nt <= 1) { work(0, height); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back(work, height * t / nt, height * (t + 1) / nt);
  for (auto &x : th) x.join();
}

int ensure_bounce(HostCache &c, size_t bytes) {
  if (c.b_bytes >= bytes) return 0;
  for (int k = 0; k < 2; ++k) {
    if (c.bin[k]) cudaFreeHost(c.bin[k]);
    if (c.bout[k]) cudaFreeHost(c.bout[k]);
    c.bin[k] = c.bout[k] = nullptr;
  }
  c.b_bytes = 0;
  for (int k = 0; k < 2; ++k) {
    HCK(cudaHostAlloc(&c.bin[k], bytes, cudaHostAllocDefault));
    HCK(cudaHostAlloc(&c.bout[k], bytes, cudaHostAllocDefault));
    if (!c.ev_bin[k]) HCK(cudaEventCreateWithFlags(&c.ev_bin[k], cudaEventDisableTiming));
    if (!c.ev_bout[k]) HCK(cudaEventCreateWithFlags(&c.ev_bout[k], cudaEventDisableTiming));
  }
  c.b_bytes = bytes;
  return 0;
}

// the thread that scatters finished panels from the pinned out-buffers back into the caller's array
struct Unpacker {
  struct Task { int k; char *dst; size_t dpitch, width, height; cudaEvent_t ev; };
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Task> q;
  bool busy[2] = {false, false}, stop = false;
  HostCache *c = nullptr;
  int dev = 0;
  std::thread th;
  void start(HostCache *cache, int device) {
    c = cache; dev = device;
    th = std::thread([this] {
      cudaSetDevice(dev);
      for (;;) {
        Task t;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [this] { return stop || !q.empty(); });
          if (q.empty()) return;
          t = q.front(); q.pop_front();
        }
        cudaEventSynchronize(t.ev);
        copy_rows_parallel(t.dst, t.dpitch, (const char *)c->bout[t.k], t.width, t.width, t.height);
        { std::lock_guard<std::mutex> lk(mu); busy[t.k] = false; }
        cv.notify_all();
      }
    });
  }
  void push(const Task &t) {
    { std::lock_guard<std::mutex> lk(mu); busy[t.k] = true; q.push_back(t); }
    cv.notify_all();
  }
  void wait_free(int k) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !busy[k]; });
  }
  void finish() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

// ---- streamed update: H2D of column panel J+1, the kernels of panel J and D2H of panel J-1
// overlap on three streams.  In the right-looking panel formulation (large.cuh) panel J only
// touches its own 256 columns (plus the n-vector w), so a panel can be finished and shipped back as
// soon as it has arrived; PCIe runs full duplex and the update costs about one H2D of the triangle.
constexpr int64_t kStreamMinN = 2048;
constexpr int64_t kXferCols = 512;  // columns per host <-> device transfer panel (measured: 256 24.7 ms, 512 24.0, 1024 27.7 at n = 16384)

template <typename T, int LAYOUT>
int run_host_update_streamed(HostCache &c, const T *Lh, int64_t ld, T *Lo, int64_t ldo, int64_t n, T *vh,
                             int32_t *status) {
  // Lh / ld: the caller's factor (read); Lo / ldo: where the updated lower triangle goes (== Lh in place)
  using namespace chub;
  const size_t es = sizeof(T);
  const int ni = (int)n;
  const size_t need_ws = large_ws_doubles(n) * sizeof(double) + 4096;
  if (c.ws_bytes < need_ws) {
    if (c.ws) HCK(cudaFree(c.ws));
    c.ws = nullptr; c.ws_bytes = 0;
    HCK(cudaMalloc(&c.ws, need_ws));
    c.ws_bytes = need_ws;
  }
  const int nJ = (ni + kBig - 1) / kBig;
  // Transfer panels are wider than the kernels' 256-column panels: a 2-D copy of 2 KB rows runs well
  // below the PCIe rate; beyond 512 columns the coarser pipeline costs more than it gains (CHUB_HOST_XFER_COLS overrides).
  int64_t xfer = kXferCols;
  if (const char *e = getenv("CHUB_HOST_XFER_COLS")) xfer = std::max<int64_t>(kBig, atoll(e) / kBig * kBig);
  const int nP = (int)((n + xfer - 1) / xfer);
  while ((int)c.ev_in.size() < nP + 1) {
    cudaEvent_t a, b;
    HCK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    HCK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    c.ev_in.push_back(a);
    c.ev_done.push_back(b);
  }
  LargeWs ws = large_ws_carve(c.ws, n);
  T *Ld = (T *)c.L;
  T *vd = (T *)c.v;
  const int64_t ldd = n;  // dense device copy
  // pageable caller memory goes through the pinned bounce buffers (CHUB_HOST_BOUNCE=0: straight copies)
  const char *be = getenv("CHUB_HOST_BOUNCE");
  const bool allow_bounce = !(be && be[0] == '0');
  const bool bounce_in = allow_bounce && !is_pinned(Lh), bounce_out = allow_bounce && !is_pinned(Lo);
  Unpacker unpacker;
  if (bounce_in || bounce_out) {
    int rc = ensure_bounce(c, (size_t)n * (size_t)std::min<int64_t>(xfer, n) * es);
    if (rc) return rc;
  }
  if (bounce_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    unpacker.start(&c, dev);
  }
  struct Fin {
    Unpacker &u;
    ~Fin() { u.finish(); }
  } fin{unpacker};
  HCK(cudaMemsetAsync(c.status, 0, sizeof(int32_t), c.st));
  HCK(cudaMemcpyAsync(vd, vh, (size_t)n * es, cudaMemcpyHostToDevice, c.st));
  large_prepare_kernel<T, LAYOUT><<<(ni + 255) / 256, 256, 0, c.st>>>(Ld, ni, ldd, vd, ws, 0, c.status);
  for (int P = 0; P < nP; ++P) {
    const int64_t j0 = (int64_t)P * xfer;
    const int64_t bw = std::min<int64_t>(xfer, n - j0);
    // rectangle rows [j0, n) x columns [j0, j0 + bw)
    size_t off_d, off_h, off_o, width, height;
    if (LAYOUT == CHUB_ROW_MAJOR) {
      off_d = (size_t)(j0 * ldd + j0); off_h = (size_t)(j0 * ld + j0); off_o = (size_t)(j0 * ldo + j0);
      width = (size_t)bw * es; height = (size_t)(n - j0);
    } else {
      off_d = (size_t)(j0 + j0 * ldd); off_h = (size_t)(j0 + j0 * ld); off_o = (size_t)(j0 + j0 * ldo);
      width = (size_t)(n - j0) * es; height = (size_t)bw;
    }
    if (bounce_in) {
      const int k = P & 1;
      if (P >= 2) HCK(cudaEventSynchronize(c.ev_bin[k]));  // the DMA of panel P-2 has drained this buffer
      copy_rows_parallel((char *)c.bin[k], width, (const char *)(Lh + off_h), (size_t)ld * es, width, height);
      HCK(cudaMemcpy2DAsync(Ld + off_d, (size_t)ldd * es, c.bin[k], width, width, height, cudaMemcpyHostToDevice,
                            c.st_in));
      HCK(cudaEventRecord(c.ev_bin[k], c.st_in));
    } else {
      HCK(cudaMemcpy2DAsync(Ld + off_d, (size_t)ldd * es, Lh + off_h, (size_t)ld * es, width, height,
                            cudaMemcpyHostToDevice, c.st_in));
    }
    HCK(cudaEventRecord(c.ev_in[P], c.st_in));
    HCK(cudaStreamWaitEvent(c.st, c.ev_in[P], 0));
    for (int J = (int)(j0 / kBig); J < nJ && (int64_t)J * kBig < j0 + bw; ++J) {
      const int64_t c0 = (int64_t)J * kBig;
      large_invdiag_kernel<T, LAYOUT><<<(kBig / 32 + 3) / 4, 128, 0, c.st>>>(Ld, ni, ldd, ws, c.status,
                                                                             (int)(c0 / 32), kBig / 32);
      large_chain_kernel<T, LAYOUT, true><<<1, kBig, 0, c.st>>>(Ld, ni, ldd, vd, ws, J, c.status);
      large_trail_kernel<T, LAYOUT, true><<<(unsigned)((n - c0 + 127) / 128), 128, 0, c.st>>>(Ld, ni, ldd, ws, J,
                                                                                             c.status);
    }
    HCK(cudaGetLastError());
    HCK(cudaEventRecord(c.ev_done[P], c.st));
    HCK(cudaStreamWaitEvent(c.st_out, c.ev_done[P], 0));
    if (bounce_out) {
      const int k = P & 1;
      unpacker.wait_free(k);  // panel P-2 has been scattered out of this buffer
      HCK(cudaMemcpy2DAsync(c.bout[k], width, Ld + off_d, (size_t)ldd * es, width, height, cudaMemcpyDeviceToHost,
                            c.st_out));
      HCK(cudaEventRecord(c.ev_bout[k], c.st_out));
      unpacker.push({k, (char *)(Lo + off_o), (size_t)ldo * es, width, height, c.ev_bout[k]});
    } else {
      HCK(cudaMemcpy2DAsync(Lo + off_o, (size_t)ldo * es, Ld + off_d, (size_t)ldd * es, width, height,
                            cudaMemcpyDeviceToHost, c.st_out));
    }
  }
  HCK(cudaMemcpyAsync(vh, vd, (size_t)n * es, cudaMemcpyDeviceToHost, c.st_out));
  HCK(cudaMemcpyAsync(status, c.status, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st_out));
  HCK(cudaStreamSynchronize(c.st_out));
  HCK(cudaStreamSynchronize(c.st));
  unpacker.wait_free(0);
  unpacker.wait_free(1);
  return 0;
}

typedef int (*dev_fn)(void *, int64_t, int64_t, int, void *, int, void *, size_t, int32_t *, void *);

// Strict upper triangle src -> dst on the host (copy mode: the reference returns a full copy of L,
// _seeger.py:272-273, so whatever sits above the diagonal travels along -- on the host, never over PCIe).
void copy_upper_host(char *dst, int64_t ldd, const char *src, int64_t lds, int64_t n, int layout, size_t es,
                     int64_t i0, int64_t i1) {
  for (int64_t i = i0; i < i1; ++i) {
    if (layout == CHUB_ROW_MAJOR) {  // row i: columns i+1 .. n-1
      if (i + 1 < n) memcpy(dst + (size_t)(i * ldd + i + 1) * es, src + (size_t)(i * lds + i + 1) * es, (size_t)(n - 1 - i) * es);
    } else {                         // column i: rows 0 .. i-1
      if (i > 0) memcpy(dst + (size_t)(i * ldd) * es, src + (size_t)(i * lds) * es, (size_t)i * es);
    }
  }
}

int run_host(dev_fn fn, bool is_update, size_t es, const void *L, int64_t ld, void *Lo, int64_t ldo, int64_t n,
             int layout, void *v, int flags, int32_t *status) {
  if (!status || n < 0 || ld < n || ldo < n) return (int)cudaErrorInvalidValue;
  *status = 0;
  if (n == 0) return 0;
  int dev = 0;
  HCK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(h_mu[dev & 63]);
  HostCache &c = h_cache[dev & 63];
  int rc = ensure(c, (size_t)n * n * es, (size_t)n * es);
  if (rc) return rc;
  if (flags & CHUB_FLAG_CHECK_DIAG) {  // _arg_validation.py:18-22, on the host copy: nothing is touched
    const char *d = (const char *)L;
    for (int64_t i = 0; i < n; ++i) {
      const char *e = d + (size_t)(i * ld + i) * es;
      const bool zero = es == 8 ? *(const double *)e == 0.0 : *(const float *)e == 0.0f;
      if (zero) { *status = CHUB_STATUS_ZERO_DIAG; return 0; }
    }
    flags &= ~CHUB_FLAG_CHECK_DIAG;
  }
  // copy mode: the strict upper triangle goes from L to Lo on host threads while the GPU works
  std::vector<std::thread> uppers;
  if (Lo != L && n > 1) {
    const int nt = n >= 4096 ? 4 : 1;
    for (int t = 0; t < nt; ++t) {
      // equal areas: the upper triangle's rows shrink (row-major) / columns grow (column-major)
      auto cut = [&](int q) -> int64_t {
        const double f = (double)q / nt;
        return layout == CHUB_ROW_MAJOR ? (int64_t)(n * (1.0 - sqrt(1.0 - f))) : (int64_t)(n * sqrt(f));
      };
      const int64_t i0 = t == 0 ? 0 : cut(t), i1 = t == nt - 1 ? n : cut(t + 1);
      uppers.emplace_back(copy_upper_host, (char *)Lo, ldo, (const char *)L, ld, n, layout, es, i0, i1);
    }
  }
  struct Joiner {
    std::vector<std::thread> &t;
    ~Joiner() { for (auto &x : t) if (x.joinable()) x.join(); }
  } joiner{uppers};
  if (is_update && n >= kStreamMinN && !getenv("CHUB_HOST_NO_STREAM")) {
    if (es == 8)
      return layout == CHUB_ROW_MAJOR
                 ? run_host_update_streamed<double, CHUB_ROW_MAJOR>(c, (const double *)L, ld, (double *)Lo, ldo, n, (double *)v, status)
                 : run_host_update_streamed<double, CHUB_COL_MAJOR>(c, (const double *)L, ld, (double *)Lo, ldo, n, (double *)v, status);
    return layout == CHUB_ROW_MAJOR
               ? run_host_update_streamed<float, CHUB_ROW_MAJOR>(c, (const float *)L, ld, (float *)Lo, ldo, n, (float *)v, status)
               : run_host_update_streamed<float, CHUB_COL_MAJOR>(c, (const float *)L, ld, (float *)Lo, ldo, n, (float *)v, status);
  }
  // device copy is dense (ld = n)
  rc = copy_tri((char *)c.L, (const char *)L, n, n, ld, layout, es, cudaMemcpyHostToDevice, c.st);
  if (rc) return rc;
  HCK(cudaMemcpyAsync(c.v, v, (size_t)n * es, cudaMemcpyHostToDevice, c.st));
  rc = fn(c.L, n, n, layout, c.v, flags, nullptr, 0, c.status, c.st);
  if (rc) {
    cudaDeviceSynchronize();  // (the message is already in chub_last_error())
    return rc;
  }
  HCK(cudaMemcpyAsync(status, c.status, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));
  HCK(cudaStreamSynchronize(c.st));
  if (*status == CHUB_STATUS_ZERO_DIAG) return 0;  // L and v untouched
  HCK(cudaMemcpyAsync(v, c.v, (size_t)n * es, cudaMemcpyDeviceToHost, c.st));
  if (*status == CHUB_STATUS_OK) {
    rc = copy_tri((char *)Lo, (const char *)c.L, n, ldo, n, layout, es, cudaMemcpyDeviceToHost, c.st);
    if (rc) return rc;
  }
  HCK(cudaStreamSynchronize(c.st));
  return 0;
}

}  // namespace

extern "C" {

int chub_update_host_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_update_f64, true, 8, L, ld, L, ld, n, layout, v, flags, status);
}
int chub_update_host_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_update_f32, true, 4, L, ld, L, ld, n, layout, v, flags, status);
}
int chub_downdate_host_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_downdate_f64, false, 8, L, ld, L, ld, n, layout, v, flags, status);
}
int chub_downdate_host_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, int32_t *status) {
  return run_host(chub_downdate_f32, false, 4, L, ld, L, ld, n, layout, v, flags, status);
}
// copy mode (overwrite_L=False, _seeger.py:272-273): L_in is only read, the result goes to L_out
int chub_update_host_copy_f64(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                              void *v, int flags, int32_t *status) {
  return run_host(chub_update_f64, true, 8, L_in, ld_in, L_out, ld_out, n, layout, v, flags, status);
}
int chub_update_host_copy_f32(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                              void *v, int flags, int32_t *status) {
  return run_host(chub_update_f32, true, 4, L_in, ld_in, L_out, ld_out, n, layout, v, flags, status);
}
int chub_downdate_host_copy_f64(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                                void *v, int flags, int32_t *status) {
  return run_host(chub_downdate_f64, false, 8, L_in, ld_in, L_out, ld_out, n, layout, v, flags, status);
}
int chub_downdate_host_copy_f32(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                                void *v, int flags, int32_t *status) {
  return run_host(chub_downdate_f32, false, 4, L_in, ld_in, L_out, ld_out, n, layout, v, flags, status);
}

int chub_host_release(void) {
  int dev = 0;
  HCK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(h_mu[dev & 63]);
  HostCache &c = h_cache[dev & 63];
  if (c.L) cudaFree(c.L);
  if (c.v) cudaFree(c.v);
  if (c.ws) cudaFree(c.ws);
  c.L = c.v = c.ws = nullptr;
  c.L_bytes = c.v_bytes = c.ws_bytes = 0;
  for (int k = 0; k < 2; ++k) {
    if (c.bin[k]) cudaFreeHost(c.bin[k]);
    if (c.bout[k]) cudaFreeHost(c.bout[k]);
    c.bin[k] = c.bout[k] = nullptr;
  }
  c.b_bytes = 0;
  return 0;
}

}  // extern "C"
