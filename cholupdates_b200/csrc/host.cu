// host.cu -- host-buffer entry points (chub_*_host_*): what a NumPy caller of the reference
// holds (_seeger.py:37-44) is staged to the device, updated there and copied back.
//
// Only the lower triangle travels: the matrix is cut into row blocks (row-major) or column
// blocks (column-major) and each block is copied as a 2-D rectangle that ends at (starts at)
// the diagonal, so PCIe carries ~n^2/2 elements each way instead of n^2.  The strict upper
// triangle of the host array is never modified.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <functional>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

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
  // staged path (run_host_staged): one pinned staging buffer holding the packed triangle blocks, one event per block
  void *stage = nullptr;
  size_t stage_bytes = 0;
  std::vector<cudaEvent_t> ev_blk;
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

// One row of a block, with non-temporal stores: the destination (staging buffer on the way in, the caller's array
// on the way out) is not read again by this core, and a regular store would first read the line it overwrites
// (3 bytes of memory traffic per byte copied instead of 2 -- host memory bandwidth is what bounds pageable callers).
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_stream_avx2(char *dst, const char *src, size_t bytes) {
  size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head > bytes) head = bytes;
  if (head) { memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
  size_t i = 0;
  for (; i + 128 <= bytes; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
  }
  for (; i + 32 <= bytes; i += 32)
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)));
  if (i < bytes) memcpy(dst + i, src + i, bytes - i);
  _mm_sfence();
}
#endif
void copy_stream(char *dst, const char *src, size_t bytes) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("CHUB_HOST_NO_NT");
  if (avx2 && bytes >= 256) { copy_stream_avx2(dst, src, bytes); return; }
#endif
  memcpy(dst, src, bytes);
}

// ---- persistent host thread pool (staged path).  Spawning threads per call costs more than a whole n = 1000
// update; the workers spin for a short while after their last task (a caller updating in a loop finds them hot)
// and then sleep on a condition variable.
class HostPool {
 public:
  static HostPool &get() {
    static HostPool *p = new HostPool(host_threads());  // (leaked on purpose: no destructor race at process exit)
    return *p;
  }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push_back(std::move(f));
      pending_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_one();
  }
  int size() const { return (int)th_.size(); }

 private:
  explicit HostPool(int n) {
    for (int t = 0; t < n; ++t) th_.emplace_back([this] { run(); });
    for (auto &x : th_) x.detach();
  }
  void run() {
    for (;;) {
      std::function<void()> f;
      if (pending_.load(std::memory_order_acquire) > 0) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!q_.empty()) {
          f = std::move(q_.front());
          q_.pop_front();
          pending_.fetch_sub(1, std::memory_order_relaxed);
        }
      }
      if (f) { f(); continue; }
      // nothing to do: spin ~200 us, then sleep
      const auto t0 = std::chrono::steady_clock::now();
      bool got = false;
      while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
        if (pending_.load(std::memory_order_acquire) > 0) { got = true; break; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      if (got) continue;
      std::unique_lock<std::mutex> lk(mu_);
      sleepers_.fetch_add(1, std::memory_order_release);
      cv_.wait(lk, [this] { return !q_.empty(); });
      sleepers_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  std::atomic<int> pending_{0}, sleepers_{0};
  std::vector<std::thread> th_;
};

// rows [0, height) of `width` bytes: strided (pitch) <-> packed, split over the host thread pool (non-temporal
// stores: the destination is not read again by this core)
void copy_rows_parallel(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t height) {
  HostPool &pool = HostPool::get();
  const int nt = (int)std::min<size_t>((size_t)pool.size() + 1, std::max<size_t>(1, height * width / (1 << 20)));
  auto work = [=](size_t r0, size_t r1) {
    for (size_t r = r0; r < r1; ++r) copy_stream(dst + r * dpitch, src + r * spitch, width);
  };
  if (nt <= 1) { work(0, height); return; }
  std::atomic<int> left{nt - 1};
  std::atomic<int> *lp = &left;
  for (int t = 1; t < nt; ++t) {
    const size_t r0 = height * t / nt, r1 = height * (t + 1) / nt;
    pool.submit([=] {
      work(r0, r1);
      lp->fetch_sub(1, std::memory_order_release);
    });
  }
  work(0, height / nt);
  while (left.load(std::memory_order_acquire) > 0) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
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
constexpr int64_t kStreamPageableMinN = 12288;  // pageable caller memory: the streamed bounce ring from here on (n = 16384:
                                               // 49-57 ms against 63-65 ms staged; n = 5000: staged 6.6-7.3 against 10-11 ms)
constexpr int64_t kStreamPinnedMinN = 4096;  // pinned caller memory: streamed from here on (else staged)
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
  // Panel boundaries.  The first panels are the tallest and nothing overlaps the first one's H2D (nor the
  // kernels that wait for it); CHUB_HOST_XFER_LEAD = number of narrow (256-column) lead panels before the
  // pipeline widens to `xfer`.  Measured at n = 16384 (gpurun r04x): 0 / 1 / 2 / 4 / 8 lead panels give
  // 25.7 / 25.3 / 26.0 / 26.0 / 25.6 ms -- the PCIe rate (~42 GB/s each way), not the fill, bounds the call.
  std::vector<int64_t> pj0;
  {
    int lead = 0;
    if (const char *e = getenv("CHUB_HOST_XFER_LEAD")) lead = std::max(0, atoi(e));
    int64_t j = 0;
    while (j < n) {
      pj0.push_back(j);
      j += ((int)pj0.size() <= lead) ? kBig : xfer;
    }
    pj0.push_back(n);
  }
  const int nP = (int)pj0.size() - 1;
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
    const int64_t j0 = pj0[P];
    const int64_t bw = std::min<int64_t>(pj0[P + 1], n) - j0;
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

// ---- staged path: any operation, any size, pageable or pinned caller memory.
// The lower triangle is cut into K blocks of about equal area (row blocks that end at the diagonal for row-major,
// column blocks that start at it for column-major).  Pool threads pack block b from the caller's (strided,
// pageable) array into a pinned staging buffer; the calling thread issues block b's DMA as soon as it is packed,
// so packing and PCIe overlap.  Then the device-resident kernels (the same entry points a CUDA-tensor caller
// reaches), then the way back: all K device -> staging copies are queued at once, pool threads scatter block b
// into the caller's array as soon as its copy has landed.  Nothing above the diagonal blocks' own little
// triangles crosses PCIe; pinned caller memory skips the staging buffer.
struct TriBlock { int64_t b0, b1; size_t off_h, off_o, off_d, width, height, stage_off; };

int run_host_staged(HostCache &c, dev_fn fn, size_t es, const void *L, int64_t ld, void *Lo, int64_t ldo, int64_t n,
                    int layout, void *v, int flags, int32_t *status) {
  HostPool &pool = HostPool::get();
  const size_t tri_bytes = (size_t)n * (size_t)(n + 1) / 2 * es;
  int K = (int)std::min<size_t>((size_t)8 * pool.size(), std::max<size_t>(1, tri_bytes / (256u << 10)));
  if (const char *e = getenv("CHUB_HOST_BLOCKS")) K = std::max(1, atoi(e));
  K = (int)std::min<int64_t>(K, (n + 31) / 32);
  std::vector<TriBlock> blk;
  size_t stage_need = 0;
  {
    int64_t prev = 0;
    for (int b = 1; b <= K && prev < n; ++b) {
      const double f = (double)b / K;
      int64_t cut = layout == CHUB_ROW_MAJOR ? (int64_t)std::llround(n * std::sqrt(f))
                                             : (int64_t)std::llround(n * (1.0 - std::sqrt(1.0 - f)));
      cut = b == K ? n : std::min<int64_t>(n, (cut + 31) / 32 * 32);
      if (cut <= prev) continue;
      TriBlock t;
      t.b0 = prev; t.b1 = cut;
      if (layout == CHUB_ROW_MAJOR) {  // rows [b0, b1) x columns [0, b1)
        t.off_h = (size_t)prev * ld; t.off_o = (size_t)prev * ldo; t.off_d = (size_t)prev * n;
        t.width = (size_t)cut * es; t.height = (size_t)(cut - prev);
      } else {                         // columns [b0, b1) x rows [b0, n)
        t.off_h = (size_t)prev * ld + prev; t.off_o = (size_t)prev * ldo + prev; t.off_d = (size_t)prev * n + prev;
        t.width = (size_t)(n - prev) * es; t.height = (size_t)(cut - prev);
      }
      t.stage_off = stage_need;
      stage_need += (t.width * t.height + 255) / 256 * 256;
      blk.push_back(t);
      prev = cut;
    }
  }
  const int nb = (int)blk.size();
  const bool pin_in = is_pinned(L), pin_out = is_pinned(Lo);
  if (!pin_in || !pin_out) {
    if (c.stage_bytes < stage_need) {
      if (c.stage) cudaFreeHost(c.stage);
      c.stage = nullptr; c.stage_bytes = 0;
      HCK(cudaHostAlloc(&c.stage, stage_need + stage_need / 8, cudaHostAllocDefault));
      c.stage_bytes = stage_need + stage_need / 8;
    }
  }
  while ((int)c.ev_blk.size() < nb) {
    cudaEvent_t e;
    HCK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c.ev_blk.push_back(e);
  }
  char *stage = (char *)c.stage;
  char *Ld = (char *)c.L;
  auto copy_rows = [](char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t height) {
    for (size_t r = 0; r < height; ++r) copy_stream(dst + r * dpitch, src + r * spitch, width);
  };

  // ---- way in
  HCK(cudaMemcpyAsync(c.v, v, (size_t)n * es, cudaMemcpyHostToDevice, c.st));
  if (pin_in) {
    for (const TriBlock &t : blk)
      HCK(cudaMemcpy2DAsync(Ld + t.off_d * es, (size_t)n * es, (const char *)L + t.off_h * es, (size_t)ld * es, t.width,
                            t.height, cudaMemcpyHostToDevice, c.st));
  } else {
    std::vector<std::atomic<int>> packed(nb);
    for (auto &x : packed) x.store(0, std::memory_order_relaxed);
    const char *Lh = (const char *)L;
    for (int b = 1; b < nb; ++b) {  // (block 0 is packed by this thread: no hand-over latency for tiny factors)
      const TriBlock t = blk[b];
      std::atomic<int> *flag = &packed[b];
      pool.submit([=] {
        copy_rows(stage + t.stage_off, t.width, Lh + t.off_h * es, (size_t)ld * es, t.width, t.height);
        flag->store(1, std::memory_order_release);
      });
    }
    copy_rows(stage + blk[0].stage_off, blk[0].width, Lh + blk[0].off_h * es, (size_t)ld * es, blk[0].width, blk[0].height);
    packed[0].store(1, std::memory_order_release);
    cudaError_t err = cudaSuccess;
    for (int b = 0; b < nb; ++b) {
      while (packed[b].load(std::memory_order_acquire) == 0) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      // (every task is waited for even after an error: the tasks reference this frame)
      if (err == cudaSuccess)
        err = cudaMemcpy2DAsync(Ld + blk[b].off_d * es, (size_t)n * es, stage + blk[b].stage_off, blk[b].width,
                                blk[b].width, blk[b].height, cudaMemcpyHostToDevice, c.st);
    }
    HCK(err);
  }

  // ---- the update / downdate, device resident
  int rc = fn(c.L, n, n, layout, c.v, flags, nullptr, 0, c.status, c.st);
  if (rc) {
    cudaDeviceSynchronize();  // (the message is already in chub_last_error())
    return rc;
  }
  HCK(cudaMemcpyAsync(status, c.status, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));

  // ---- way out.  Into the staging buffer the copies start at once (nothing reaches the caller's array before the
  // status is known); straight into pinned caller memory only after the status has been read.
  if (!pin_out) {
    for (int b = 0; b < nb; ++b) {
      HCK(cudaMemcpy2DAsync(stage + blk[b].stage_off, blk[b].width, Ld + blk[b].off_d * es, (size_t)n * es, blk[b].width,
                            blk[b].height, cudaMemcpyDeviceToHost, c.st));
      HCK(cudaEventRecord(c.ev_blk[b], c.st));
    }
  }
  if (!pin_out) HCK(cudaEventSynchronize(c.ev_blk[0]));
  else HCK(cudaStreamSynchronize(c.st));
  // (status was queued before the first block: it is there once block 0 has landed / the stream is idle)
  if (*status == CHUB_STATUS_ZERO_DIAG) { HCK(cudaStreamSynchronize(c.st)); return 0; }  // L and v untouched
  if (*status != CHUB_STATUS_OK) {  // not positive definite: v holds p, L stays as it was
    HCK(cudaMemcpyAsync(v, c.v, (size_t)n * es, cudaMemcpyDeviceToHost, c.st));
    HCK(cudaStreamSynchronize(c.st));
    return 0;
  }
  HCK(cudaMemcpyAsync(v, c.v, (size_t)n * es, cudaMemcpyDeviceToHost, c.st));
  if (pin_out) {
    for (const TriBlock &t : blk)
      HCK(cudaMemcpy2DAsync((char *)Lo + t.off_o * es, (size_t)ldo * es, Ld + t.off_d * es, (size_t)n * es, t.width,
                            t.height, cudaMemcpyDeviceToHost, c.st));
    HCK(cudaStreamSynchronize(c.st));
    return 0;
  }
  {
    std::atomic<int> left{nb - 1};
    std::atomic<int> bad{0};
    char *Loh = (char *)Lo;
    for (int b = 1; b < nb; ++b) {
      const TriBlock t = blk[b];
      cudaEvent_t ev = c.ev_blk[b];
      std::atomic<int> *lp = &left, *bp = &bad;
      pool.submit([=] {
        if (cudaEventSynchronize(ev) != cudaSuccess) bp->store(1);
        else copy_rows(Loh + t.off_o * es, (size_t)ldo * es, stage + t.stage_off, t.width, t.width, t.height);
        lp->fetch_sub(1, std::memory_order_release);
      });
    }
    copy_rows(Loh + blk[0].off_o * es, (size_t)ldo * es, stage + blk[0].stage_off, blk[0].width, blk[0].width, blk[0].height);
    while (left.load(std::memory_order_acquire) > 0) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    HCK(cudaStreamSynchronize(c.st));
    if (bad.load()) HCK(cudaErrorUnknown);
  }
  return 0;
}


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
  // Which path: the streamed update overlaps the two PCIe directions (about ONE triangle transfer per update) but
  // runs the launch-per-panel kernels and wants pinned memory; the staged path takes anything and uses the
  // persistent kernels.  CHUB_HOST_PATH = streamed | staged | direct overrides (direct: plain 2-D copies, A/B).
  const char *hp = getenv("CHUB_HOST_PATH");
  bool streamed = is_update && ((n >= kStreamPinnedMinN && is_pinned(L) && is_pinned(Lo)) || n >= kStreamPageableMinN);
  if (hp && !strcmp(hp, "streamed")) streamed = is_update && n >= kStreamMinN;
  if (hp && (!strcmp(hp, "staged") || !strcmp(hp, "direct"))) streamed = false;
  if (getenv("CHUB_HOST_NO_STREAM")) streamed = false;
  // (the staged path holds the packed triangle in pinned memory: beyond CHUB_HOST_STAGE_MAX_MB, default 6 GB
  // = n ~ 38000 in fp64, pageable callers go back to the streamed bounce ring (updates) or plain 2-D copies)
  size_t stage_max = (size_t)6144 << 20;
  if (const char *e = getenv("CHUB_HOST_STAGE_MAX_MB")) stage_max = (size_t)std::max(0LL, atoll(e)) << 20;
  const bool stage_fits = (size_t)n * (size_t)(n + 1) / 2 * es <= stage_max || (is_pinned(L) && is_pinned(Lo));
  if (!stage_fits && !streamed && is_update && n >= kStreamMinN && !getenv("CHUB_HOST_NO_STREAM")) streamed = true;
  if (!streamed && stage_fits && !(hp && !strcmp(hp, "direct")))
    return run_host_staged(c, fn, es, L, ld, Lo, ldo, n, layout, v, flags, status);
  if (streamed) {
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
  if (c.stage) cudaFreeHost(c.stage);
  c.stage = nullptr;
  c.stage_bytes = 0;
  return 0;
}

}  // extern "C"
