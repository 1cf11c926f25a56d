// api.cu -- the C ABI declared in include/cholupdates_b200.h.
//
// Replaces the reference's inner kernels _seeger_impl_cython.update / .downdate
// (src/cholupdates/rank_1/_seeger_impl_cython.pyx:22-155, :160-330) with sm_100a kernels.
// No argument validation here (the reference kernels do none either, .pyx:30-31); the host
// front-end (cholupdates_b200/rank_1/_seeger.py) reproduces the reference's checks.

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "dataflow.cuh"
#include "large.cuh"
#include "rowcyclic.cuh"
#include "small.cuh"

using namespace chub;

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_path{0};

int fail(cudaError_t e, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
#define CK(call)                                        \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return fail(e_, #call);      \
  } while (0)
#define LAUNCHED()                                      \
  do {                                                  \
    g_launches.fetch_add(1, std::memory_order_relaxed); \
    CK(cudaGetLastError());                             \
  } while (0)

// ---- internal per-device workspace (used when the caller passes workspace == NULL)
struct DevCache {
  void *ws = nullptr;
  size_t ws_bytes = 0;
};
std::mutex g_mu;
DevCache g_cache[64];

int internal_ws(size_t need, void **out) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  DevCache &c = g_cache[dev & 63];
  if (c.ws_bytes < need) {
    if (c.ws) CK(cudaFree(c.ws));
    c.ws = nullptr;
    c.ws_bytes = 0;
    CK(cudaMalloc(&c.ws, need));
    c.ws_bytes = need;
  }
  *out = c.ws;
  return 0;
}

template <typename T> constexpr int dtype_of() { return sizeof(T) == 8 ? CHUB_DTYPE_F64 : CHUB_DTYPE_F32; }

// ---- single-CTA kernels (also the batched path)
template <typename T, bool UPDATE>
int launch_small(T *L, int64_t n, int64_t ld, int64_t stride_L, int layout, T *v, int64_t stride_v,
                 int64_t batch, int flags, int32_t *status, cudaStream_t st) {
  SmallArgs<T> a{L, v, status, (int)n, ld, stride_L, stride_v, flags};
  // Threads per factor.  A single factor wants the whole CTA on its trailing rows; a large batch is
  // throughput-bound, and narrow CTAs (more factors resident per SM, nobody idling behind the one
  // warp that walks the diagonal block) hide the sequential chain behind other factors.  Measured at
  // 8192 x n=256 fp64: update 1.23 ms (128 threads) / 1.18 (64) / 0.96 (32); downdate 1.76 / 1.67 / 1.84.
  // (Tried and dropped: register-direct trailing rows, 4 lanes per row -- 1.33-1.54 ms.)
  int nt = kSmallNT;
  if (const char *e = getenv("CHUB_SMALL_NT")) nt = atoi(e);
  else if (batch >= 2048) nt = UPDATE ? kSmallNTBatched : kSmallNTBatchedDd;
  auto go = [&](auto kern, int threads) -> int {
    const size_t smem = UPDATE ? small_update_smem<T>((int)n, layout, threads) : small_downdate_smem<T>((int)n, layout, threads);
    if (smem > 48 * 1024)
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)batch, threads, smem, st>>>(a);
    LAUNCHED();
    return 0;
  };
#define CHUB_SMALL_GO(NT)                                                                        \
  do {                                                                                           \
    if (UPDATE) {                                                                                \
      if (layout == CHUB_ROW_MAJOR) return go(small_update_kernel<T, CHUB_ROW_MAJOR, NT>, NT);   \
      return go(small_update_kernel<T, CHUB_COL_MAJOR, NT>, NT);                                 \
    }                                                                                            \
    if (layout == CHUB_ROW_MAJOR) return go(small_downdate_kernel<T, CHUB_ROW_MAJOR, NT>, NT);   \
    return go(small_downdate_kernel<T, CHUB_COL_MAJOR, NT>, NT);                                 \
  } while (0)
  if (nt == 32) CHUB_SMALL_GO(32);
  if (nt == 64) CHUB_SMALL_GO(64);
  CHUB_SMALL_GO(kSmallNT);
#undef CHUB_SMALL_GO
}

// ---- launch-per-panel path
template <typename T, int LAYOUT, bool UPDATE>
int launch_panel(T *L, int64_t n, int64_t ld, T *v, int flags, const LargeWs &ws, int32_t *status,
                 cudaStream_t st) {
  const int ni = (int)n;
  large_prepare_kernel<T, LAYOUT><<<(ni + 255) / 256, 256, 0, st>>>(L, ni, ld, v, ws, flags, status);
  LAUNCHED();
  const int nb = (ni + 31) / 32;
  large_invdiag_kernel<T, LAYOUT><<<(nb + 3) / 4, 128, 0, st>>>(L, ni, ld, ws, status);
  LAUNCHED();
  const int nJ = (ni + kBig - 1) / kBig;
  for (int J = 0; J < nJ; ++J) {
    large_chain_kernel<T, LAYOUT, UPDATE><<<1, kBig, 0, st>>>(L, ni, ld, v, ws, J, status);
    LAUNCHED();
    const int rbeg = J * kBig + (UPDATE ? 0 : kBig);
    if (rbeg < ni) {
      large_trail_kernel<T, LAYOUT, UPDATE><<<(ni - rbeg + 127) / 128, 128, 0, st>>>(L, ni, ld, ws, J, status);
      LAUNCHED();
    }
  }
  if (!UPDATE) {
    large_dd_coef_kernel<T, LAYOUT><<<1, 1024, 0, st>>>(L, ni, ld, v, ws, status);
    LAUNCHED();
    large_dd_sweep_kernel<T, LAYOUT><<<(ni + 127) / 128, 128, 0, st>>>(L, ni, ld, ws, status);
    LAUNCHED();
  }
  return 0;
}

constexpr int64_t kSmallCutoff = 95;    // n <= cutoff: one CTA does the whole factor (measured: the
                                        // persistent kernel wins from n = 96, tools/cutoff_probe.py)
constexpr int64_t kSmallFallback = 256;  // shapes the persistent kernel cannot take: one CTA up to here

template <typename T, bool UPDATE>
int run_single(void *Lv, int64_t n, int64_t ld, int layout, void *vv, int flags, void *workspace,
               size_t workspace_bytes, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n < 0 || ld < n || (layout != CHUB_ROW_MAJOR && layout != CHUB_COL_MAJOR) || !status) {
    snprintf(g_err, sizeof(g_err), "invalid argument (n=%lld ld=%lld layout=%d)", (long long)n,
             (long long)ld, layout);
    return (int)cudaErrorInvalidValue;
  }
  CK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  if (n == 0) return 0;
  T *L = (T *)Lv;
  T *v = (T *)vv;
  int path = g_path.load();
  if (path == 0) path = n <= kSmallCutoff ? 1 : 3;
  if (path == 1 && n > kSmallMaxN) path = 2;
  // (row-major beyond the persistent kernel's capacity: blocked driver; column-major shapes the native
  // kernel does not take: row-major scratch -- both decided below)
  if (path == 3 && !dataflow_supported<T>(n, ld, layout, L) && !blocked_supported<T>(n, ld, L, layout) &&
      !(layout == CHUB_COL_MAJOR && n > kSmallFallback))
    path = (g_path.load() == 0 && n <= kSmallFallback) ? 1 : 2;
  if (path == 1) return launch_small<T, UPDATE>(L, n, ld, 0, layout, v, 0, 1, flags, status, st);

  const size_t need = chub_workspace_bytes(dtype_of<T>(), n);
  if (!workspace) {
    int rc = internal_ws(need, &workspace);
    if (rc) return rc;
  } else if (workspace_bytes < need) {
    snprintf(g_err, sizeof(g_err), "workspace too small: %zu < %zu", workspace_bytes, need);
    return (int)cudaErrorInvalidValue;
  }
  LargeWs ws = large_ws_carve(workspace, n);
  if (path == 3 && blocked_supported<T>(n, ld, L, layout) &&
      (layout == CHUB_ROW_MAJOR || !dataflow_native<T>(n, ld, layout, L))) {
    // beyond the persistent kernel's row capacity (row-major: one launch takes n <= 16464; column-major: the
    // two-warp workers take n <= 28224): diagonal blocks + rectangles (launch_blocked)
    int64_t nl = 0;
    int rc = layout == CHUB_ROW_MAJOR ? launch_blocked<T, CHUB_ROW_MAJOR, UPDATE>(L, n, ld, v, flags, ws, status, st, &nl)
                                      : launch_blocked<T, CHUB_COL_MAJOR, UPDATE>(L, n, ld, v, flags, ws, status, st, &nl);
    g_launches.fetch_add(nl, std::memory_order_relaxed);
    if (rc == 0) return 0;
    return fail((cudaError_t)rc, "blocked dataflow launch");
  }
  if (path == 3 && layout == CHUB_COL_MAJOR && !dataflow_native<T>(n, ld, layout, L) &&
      !getenv("CHUB_COL_MAJOR_NATIVE")) {
    // Column-major shapes the native tensor-map kernel (df_kernel_cm) does not take (beyond its row cap,
    // odd leading dimension): go through a row-major scratch copy of the lower triangle -- two extra
    // coalesced triangle passes cost less than the older column-major kernel (df_kernel_v2).
    constexpr int64_t kEpc = 16 / sizeof(T);
    const int64_t ldr = (n + kEpc - 1) / kEpc * kEpc;  // 16-byte aligned rows
    T *tmp = nullptr;
    {  // keep the scratch cached in the device's default pool between calls
      static std::atomic<unsigned long long> pool_tuned{0};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!((pool_tuned.load() >> (dev & 63)) & 1ull)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
          unsigned long long thr = ~0ull;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        pool_tuned.fetch_or(1ull << (dev & 63));
      }
    }
    if ((dataflow_supported<T>(n, ldr, CHUB_ROW_MAJOR, (const T *)nullptr) ||
         blocked_supported<T>(n, ldr, (const T *)nullptr)) &&
        cudaMallocAsync((void **)&tmp, (size_t)n * ldr * sizeof(T), st) == cudaSuccess) {
      const int nt = (int)((n + 31) / 32);
      const unsigned tiles = (unsigned)((int64_t)nt * (nt + 1) / 2);
      tri_transpose_kernel<T, true><<<tiles, 256, 0, st>>>(tmp, ldr, L, ld, (int)n);
      LAUNCHED();
      int64_t nl = 0;
      int rc = blocked_supported<T>(n, ldr, tmp)
                   ? launch_blocked<T, CHUB_ROW_MAJOR, UPDATE>(tmp, n, ldr, v, flags, ws, status, st, &nl)
                   : launch_dataflow<T, CHUB_ROW_MAJOR, UPDATE>(tmp, n, ldr, v, flags, ws, status, st, &nl);
      g_launches.fetch_add(nl, std::memory_order_relaxed);
      if (rc == 0) {
        // (a failed check_diag / not-PD call leaves tmp == L, so copying back is harmless)
        tri_transpose_kernel<T, false><<<tiles, 256, 0, st>>>(L, ld, tmp, ldr, (int)n);
        LAUNCHED();
        CK(cudaFreeAsync(tmp, st));
        return 0;
      }
      cudaFreeAsync(tmp, st);
      if (rc != (int)cudaErrorNotSupported && rc != (int)cudaErrorCooperativeLaunchTooLarge)
        return fail((cudaError_t)rc, "dataflow launch");
      cudaGetLastError();
    } else {
      cudaGetLastError();
    }
  }
  if (path == 3) {
    int64_t nl = 0;
    int rc = layout == CHUB_ROW_MAJOR
                 ? launch_dataflow<T, CHUB_ROW_MAJOR, UPDATE>(L, n, ld, v, flags, ws, status, st, &nl)
                 : launch_dataflow<T, CHUB_COL_MAJOR, UPDATE>(L, n, ld, v, flags, ws, status, st, &nl);
    g_launches.fetch_add(nl, std::memory_order_relaxed);
    if (rc == 0) return 0;
    if (rc != (int)cudaErrorNotSupported && rc != (int)cudaErrorCooperativeLaunchTooLarge)
      return fail((cudaError_t)rc, "dataflow launch");
    cudaGetLastError();  // the cooperative launch was refused: use the launch-per-panel kernels
  }
  return layout == CHUB_ROW_MAJOR
             ? launch_panel<T, CHUB_ROW_MAJOR, UPDATE>(L, n, ld, v, flags, ws, status, st)
             : launch_panel<T, CHUB_COL_MAJOR, UPDATE>(L, n, ld, v, flags, ws, status, st);
}

template <typename T, bool UPDATE>
int run_batched(void *Lv, int64_t n, int64_t ld, int64_t stride_L, int layout, void *vv,
                int64_t stride_v, int64_t batch, int flags, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n < 0 || ld < n || batch < 0 || !status ||
      (layout != CHUB_ROW_MAJOR && layout != CHUB_COL_MAJOR)) {
    snprintf(g_err, sizeof(g_err), "invalid argument");
    return (int)cudaErrorInvalidValue;
  }
  if (batch == 0) return 0;
  CK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)batch, st));
  if (n == 0) return 0;
  if (n <= kSmallMaxN)
    return launch_small<T, UPDATE>((T *)Lv, n, ld, stride_L, layout, (T *)vv, stride_v, batch, flags,
                                   status, st);
  // very large factors in a batch: one after the other through the single-factor path
  for (int64_t b = 0; b < batch; ++b) {
    int rc = run_single<T, UPDATE>((T *)Lv + b * stride_L, n, ld, layout, (T *)vv + b * stride_v,
                                   flags, nullptr, 0, status + b, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

// ---- block-row-cyclic single factor (per-panel pieces; the host driver does the broadcast)
namespace {
template <typename T>
int rc_prepare(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v, void *workspace,
               int flags, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LargeWs ws = large_ws_carve(workspace, n);
  CK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  rc_prepare_kernel<T><<<(unsigned)((m_loc + 255) / 256), 256, 0, st>>>((const T *)A, (int)m_loc, (int)n, ld, G, rank,
                                                                        (const T *)v, ws, flags, status);
  LAUNCHED();
  return 0;
}
template <typename T>
int rc_panel_owner(void *A, int64_t n, int64_t ld, int G, int rank, int J, void *v_out, void *workspace,
                   double *payload, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LargeWs ws = large_ws_carve(workspace, n);
  // the diagonal block of panel J is local block q = J / G; shift the base so that the panel
  // kernels can address it with GLOBAL row indices
  const int64_t q = J / G;
  (void)rank;
  T *Lsh = (T *)A + (q * kRcBlock - (int64_t)J * kRcBlock) * ld;
  large_invdiag_kernel<T, CHUB_ROW_MAJOR><<<(kBig / 32 + 3) / 4, 128, 0, st>>>(Lsh, (int)n, ld, ws, status,
                                                                                J * (kBig / 32), kBig / 32);
  LAUNCHED();
  large_chain_kernel<T, CHUB_ROW_MAJOR, true><<<1, kBig, 0, st>>>(Lsh, (int)n, ld, (T *)v_out, ws, J, status);
  LAUNCHED();
  rc_pack_kernel<<<1, kBig, 0, st>>>(ws, J, (int)((n + 31) & ~31LL), payload, status);
  LAUNCHED();
  return 0;
}
template <typename T>
int rc_panel_apply(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J, int is_owner,
                   void *workspace, const double *payload, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LargeWs ws = large_ws_carve(workspace, n);
  if (!is_owner) {
    rc_unpack_kernel<<<1, kBig, 0, st>>>(ws, J, (int)((n + 31) & ~31LL), payload, status);
    LAUNCHED();
  }
  if (m_loc > 0) {
    rc_trail_kernel<T><<<(unsigned)((m_loc + 127) / 128), 128, 0, st>>>((T *)A, (int)m_loc, (int)n, ld, G, rank, ws, J,
                                                                       status);
    LAUNCHED();
  }
  return 0;
}
// ---- distributed downdate: forward substitution panel by panel, then a local sweep
template <typename T>
int rc_dd_panel_owner(void *A, int64_t n, int64_t ld, int G, int rank, int J, void *workspace, double *payload,
                      int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LargeWs ws = large_ws_carve(workspace, n);
  (void)rank;
  T *Lsh = (T *)A + ((int64_t)(J / G) - (int64_t)J) * kRcBlock * ld;
  large_invdiag_kernel<T, CHUB_ROW_MAJOR><<<(kBig / 32 + 3) / 4, 128, 0, st>>>(Lsh, (int)n, ld, ws, status,
                                                                                J * (kBig / 32), kBig / 32);
  LAUNCHED();
  large_chain_kernel<T, CHUB_ROW_MAJOR, false><<<1, kBig, 0, st>>>(Lsh, (int)n, ld, (T *)nullptr, ws, J, status);
  LAUNCHED();
  rc_pack_dd_kernel<T><<<1, kBig, 0, st>>>(Lsh, (int)n, ld, ws, J, payload, status);
  LAUNCHED();
  return 0;
}
template <typename T>
int rc_dd_panel_apply(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J, void *workspace,
                      const double *payload, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LargeWs ws = large_ws_carve(workspace, n);
  rc_unpack_dd_kernel<<<1, kBig, 0, st>>>(ws, J, (int)((n + 31) & ~31LL), payload, status);  // (the owner too: it sets the signs)
  LAUNCHED();
  if (m_loc > 0 && (int64_t)(J + 1) * kBig < n) {
    rc_trail_kernel<T, false><<<(unsigned)((m_loc + 127) / 128), 128, 0, st>>>((T *)A, (int)m_loc, (int)n, ld, G, rank,
                                                                              ws, J, status);
    LAUNCHED();
  }
  return 0;
}
// rho^2 = 1 - p.p, the status word (CHUB_STATUS_NOT_PD leaves A untouched and v <- p), coefficients and
// rotg's z into v -- computed redundantly (and identically) on every rank -- then the local sweep.
template <typename T>
int rc_dd_finish(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v, void *workspace,
                 int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LargeWs ws = large_ws_carve(workspace, n);
  large_dd_coef_kernel<T, CHUB_ROW_MAJOR><<<1, 1024, 0, st>>>((const T *)nullptr, (int)n, ld, (T *)v, ws, status);
  LAUNCHED();
  if (m_loc > 0) {
    rc_dd_sweep_kernel<T><<<(unsigned)((m_loc + 127) / 128), 128, 0, st>>>((T *)A, (int)m_loc, (int)n, ld, G, rank, ws,
                                                                          status);
    LAUNCHED();
  }
  return 0;
}

// ---- wavefront over K successive updates of a block-row-cyclic factor (rowcyclic.cuh)
template <typename T>
int rcw_prepare(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, const void *V, int64_t K,
                int64_t ldv, void *workspace, int flags, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0 || K <= 0 || G <= 0 || K > 65535) {
    snprintf(g_err, sizeof(g_err), "invalid argument (n=%lld K=%lld G=%d)", (long long)n, (long long)K, G);
    return (int)cudaErrorInvalidValue;
  }
  RcWave wv = rcw_carve(workspace, n, K, G);
  CK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  dim3 grid((unsigned)((m_loc > 0 ? m_loc + 255 : 256) / 256), (unsigned)K);
  rcw_prepare_kernel<T><<<grid, 256, 0, st>>>((const T *)A, (int)m_loc, (int)n, ld, G, rank, (const T *)V, ldv, wv,
                                               flags, status);
  LAUNCHED();
  return 0;
}
// Peer-memory exchange arguments.  The slot stride is a property of the mailbox (rcw_mail_tasks), not of
// the call; the watchdog is generous because ranks may enter a call seconds apart (CHUB_P2P_WATCHDOG_MS).
inline RcwP2P rcw_p2p_args(const void *peer_table, void *mailbox, int64_t epoch, int64_t n, int G) {
  static long long cycles = 0;
  if (!cycles) {
    long long ms = 20000;
    if (const char *e = getenv("CHUB_P2P_WATCHDOG_MS")) ms = atoll(e) > 0 ? atoll(e) : ms;
    cycles = ms * 2000000LL;  // ~2 GHz
  }
  return RcwP2P{(double *const *)peer_table, (double *)mailbox, (long long)epoch, rcw_mail_tasks(n, G), cycles};
}
// tasks of anti-diagonal d: panels Jlo..Jhi (update u = d - J)
inline void rcw_span(int64_t n, int64_t K, int64_t d, int *Jlo, int *Jhi) {  // K = updates, or groups when fused
  const int64_t nJ = rcw_panels(n);
  *Jlo = (int)(d - K + 1 > 0 ? d - K + 1 : 0);
  *Jhi = (int)(d < nJ - 1 ? d : nJ - 1);
}
template <typename T>
int rcw_owner(void *A, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d, void *workspace,
              double *pay_mine, const void *peer_table, void *mailbox, int64_t epoch, int solve_only, int fuse,
              int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (fuse < 1 || fuse > kRcwMaxFuse || (solve_only && fuse != 1)) return (int)cudaErrorInvalidValue;
  RcWave wv = rcw_carve(workspace, n, K, G);
  wv.F = fuse;
  RcwP2P pp = rcw_p2p_args(peer_table, mailbox, epoch, n, G);
  int Jlo, Jhi;
  rcw_span(n, rcw_groups(K, fuse), d, &Jlo, &Jhi);
  const int Jfirst = Jlo + (((rank - Jlo) % G) + G) % G;
  const int count = Jfirst > Jhi ? 0 : (Jhi - Jfirst) / G + 1;
  if (count == 0 && !peer_table) return 0;
  // (P2P mode: a rank without a task at this step still raises its flags -- one CTA)
  auto kern = solve_only ? rcw_owner_kernel<T, false> : rcw_owner_kernel<T, true>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRcwOwnerSmem));
  kern<<<count > 0 ? count : 1, kBig, kRcwOwnerSmem, st>>>((const T *)A, (int)n, ld, G, rank, wv, (int)d, Jfirst,
                                                           pay_mine, pp, count, status);
  LAUNCHED();
  return 0;
}
template <typename T>
int rcw_apply(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
              void *workspace, const double *pay_all, const void *peer_table, void *mailbox, void *V_out,
              int64_t ldv, int64_t epoch, int rows_mode, void *solve_ws, int fuse, int32_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (fuse < 1 || fuse > kRcwMaxFuse || (solve_ws && fuse != 1)) return (int)cudaErrorInvalidValue;
  RcWave wv = rcw_carve(workspace, n, K, G);
  wv.F = fuse;
  RcwP2P pp = rcw_p2p_args(peer_table, mailbox, epoch, n, G);
  int Jlo, Jhi;
  rcw_span(n, rcw_groups(K, fuse), d, &Jlo, &Jhi);
  if (Jhi < Jlo) return 0;
  int nt = kRcwTrailNT;
  if (const char *e = getenv("CHUB_RCW_NT")) nt = atoi(e);
  if (nt != 32 && nt != 64) nt = 128;
  if (rows_mode < 0 || rows_mode > 2) return (int)cudaErrorInvalidValue;
  // bulk of the look-ahead split on the TMA ring (rcw_bulk_tma_kernel) when the local rows are 16-byte aligned
  // (opt-in, CHUB_RCW_BULK_TMA=1: measured slower than the staged kernel so far -- 0.63 vs 0.54 ms per update at
  // n = 16384, K = 32 -- because an item is only 8 tiles long and the ring does not yet prefetch across items)
  if (rows_mode == 2 && !solve_ws && fuse == 1 && m_loc > 0 && getenv("CHUB_RCW_BULK_TMA") &&
      (ld % (16 / sizeof(T))) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && get_encode_tiled()) {
    alignas(64) CUtensorMap tmapA;
    if (make_tmap<T>(&tmapA, (T *)A, n, ld, 8, m_loc)) {
      static int sms = 0;
      if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      }
      const char *bt = getenv("CHUB_RCW_BULK_TMA");
      auto kern = (bt && bt[0] == '1') ? rcw_bulk_tma_kernel<T> : rcw_bulk_tma2_kernel<T>;  // 1: per-item pipeline (A/B)
      const size_t smem = rect_apply_rm_smem<T>();
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int ntasks = Jhi - Jlo + 1;
      const int64_t items = (int64_t)ntasks * ((m_loc + 7) / 8);  // upper bound
      // The bulk runs next to the owner + head kernels of the NEXT step (high-priority stream).  Its CTAs are
      // persistent and fill an SM's shared memory, so a few SMs are left free for those kernels -- otherwise
      // they would queue behind the whole bulk and the look-ahead would be lost.
      int free_sms = 12;
      if (const char *e = getenv("CHUB_RCW_FREE_SMS")) free_sms = atoi(e);
      const int cap = std::max(1, sms - std::max(0, free_sms));
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(cap, (items + kRectWarps - 1) / kRectWarps));
      kern<<<grid, kRectWarps * 32, smem, st>>>((int)m_loc, (int)n, G, rank, wv, (int)d, Jlo, ntasks,
                                                rcw_max_tasks(n, K, G), pay_all, (T *)V_out, ldv, pp, status, tmapA);
      LAUNCHED();
      return 0;
    }
  }
  // solve_ws != NULL: forward substitution only (downdate); p and the diagonal signs go into its p / dn
  double *p_out = nullptr, *dn_out = nullptr;
  if (solve_ws) {
    LargeWs lw = large_ws_carve(solve_ws, n);
    p_out = lw.p;
    dn_out = lw.dn;
  }
  // head of the look-ahead split: only the CTAs of the two 256-row blocks J and J+1 of every task
  const unsigned gx = rows_mode == 1 ? (unsigned)(2 * kRcBlock / nt) : (unsigned)((m_loc > 0 ? m_loc + nt - 1 : nt) / nt);
  dim3 grid(gx, (unsigned)(Jhi - Jlo + 1));
#define CHUB_TRAIL_GO(NTV, UPD, FV)                                                                                  \
  rcw_trail_kernel<T, NTV, UPD, FV><<<grid, NTV, 0, st>>>((T *)A, (int)m_loc, (int)n, ld, G, rank, wv, (int)d, Jlo,  \
                                                          rcw_max_tasks(n, K, G), pay_all, (T *)V_out, ldv, pp,      \
                                                          rows_mode, p_out, dn_out, status)
  if (solve_ws) {
    if (nt == 32) CHUB_TRAIL_GO(32, false, 1);
    else if (nt == 64) CHUB_TRAIL_GO(64, false, 1);
    else CHUB_TRAIL_GO(128, false, 1);
  } else if (fuse == 2) {
    if (nt == 32) CHUB_TRAIL_GO(32, true, 2);
    else if (nt == 64) CHUB_TRAIL_GO(64, true, 2);
    else CHUB_TRAIL_GO(128, true, 2);
  } else {
    if (nt == 32) CHUB_TRAIL_GO(32, true, 1);
    else if (nt == 64) CHUB_TRAIL_GO(64, true, 1);
    else CHUB_TRAIL_GO(128, true, 1);
  }
#undef CHUB_TRAIL_GO
  LAUNCHED();
  return 0;
}
}  // namespace

extern "C" {

int chub_version(void) { return 120; }  // 120: resident factors, chub_*_many; 110: chub_rcw_* (wavefront), chub_rc_dd_*, chub_p2p_*
const char *chub_last_error(void) { return g_err; }
// (internal, not in the header) host.cu reports its failures through the same buffer
void chub_internal_set_error(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }

int chub_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int chub_device_info(int *sm_count, int64_t *l2_bytes, int *cc_major, int *cc_minor) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (l2_bytes) *l2_bytes = p.l2CacheSize;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

size_t chub_workspace_bytes(int dtype, int64_t n) {
  (void)dtype;
  if (n < 0) n = 0;
  return large_ws_doubles(n) * sizeof(double) + dataflow_flag_bytes(n);
}

int chub_set_path(int path) { return g_path.exchange(path); }

// Debug: copy the dataflow kernel's cycle counters (CHUB_PROFILE=1) of the internal or given
// workspace to the host: out[cta * 16 + slot].
int chub_debug_profile(void *workspace, int64_t n, long long *out, int ctas) {
  if (!workspace) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    workspace = g_cache[dev & 63].ws;
    if (!workspace) return (int)cudaErrorInvalidValue;
  }
  LargeWs ws = large_ws_carve(workspace, n);
  CK(cudaDeviceSynchronize());
  // ctas < 0: copy -ctas raw counters (cycle counters followed by the event trace)
  const size_t count = ctas >= 0 ? (size_t)ctas * 16 : (size_t)(-ctas);
  CK(cudaMemcpy(out, reinterpret_cast<char *>(ws.flags) + 256, count * sizeof(long long),
                cudaMemcpyDeviceToHost));
  return 0;
}
int64_t chub_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int chub_rc_payload_doubles(void) { return kRcPayload; }
int chub_rc_block_rows(void) { return kRcBlock; }
int chub_rc_prepare_f64(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v, void *workspace,
                        int flags, int32_t *status, void *stream) {
  return rc_prepare<double>(A, m_loc, n, ld, G, rank, v, workspace, flags, status, stream);
}
int chub_rc_prepare_f32(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v, void *workspace,
                        int flags, int32_t *status, void *stream) {
  return rc_prepare<float>(A, m_loc, n, ld, G, rank, v, workspace, flags, status, stream);
}
int chub_rc_panel_owner_f64(void *A, int64_t n, int64_t ld, int G, int rank, int J, void *v_out, void *workspace,
                            double *payload, int32_t *status, void *stream) {
  return rc_panel_owner<double>(A, n, ld, G, rank, J, v_out, workspace, payload, status, stream);
}
int chub_rc_panel_owner_f32(void *A, int64_t n, int64_t ld, int G, int rank, int J, void *v_out, void *workspace,
                            double *payload, int32_t *status, void *stream) {
  return rc_panel_owner<float>(A, n, ld, G, rank, J, v_out, workspace, payload, status, stream);
}
int chub_rc_panel_apply_f64(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J, int is_owner,
                            void *workspace, const double *payload, int32_t *status, void *stream) {
  return rc_panel_apply<double>(A, m_loc, n, ld, G, rank, J, is_owner, workspace, payload, status, stream);
}
int chub_rc_panel_apply_f32(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J, int is_owner,
                            void *workspace, const double *payload, int32_t *status, void *stream) {
  return rc_panel_apply<float>(A, m_loc, n, ld, G, rank, J, is_owner, workspace, payload, status, stream);
}

// ---- peer-memory plumbing for the P2P exchange (cudaIpc: one process per GPU)
size_t chub_rcw_mailbox_bytes(int64_t n, int64_t K, int G) {
  if (n < 0) n = 0;
  if (K < 1) K = 1;
  return rcw_mailbox_doubles(n, K, G > 0 ? G : 1) * sizeof(double);
}
int chub_p2p_alloc(size_t bytes, void **ptr) {
  if (!ptr) return (int)cudaErrorInvalidValue;
  CK(cudaMalloc(ptr, bytes));
  CK(cudaMemset(*ptr, 0, bytes));
  CK(cudaDeviceSynchronize());
  return 0;
}
int chub_p2p_free(void *ptr) {
  if (ptr) CK(cudaFree(ptr));
  return 0;
}
int chub_p2p_export(void *ptr, void *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
int chub_p2p_import(const void *handle64, void **peer_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int chub_p2p_close(void *peer_ptr) {
  if (peer_ptr) CK(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

int chub_rcw_payload_doubles(void) { return kRcwSlot; }
int chub_rcw_max_fuse(void) { return kRcwMaxFuse; }
int chub_rcw_max_tasks(int64_t n, int64_t K, int G) { return rcw_max_tasks(n, K, G > 0 ? G : 1); }
size_t chub_rcw_workspace_bytes(int64_t n, int64_t K, int G) {
  if (n < 0) n = 0;
  if (K < 0) K = 0;
  return rcw_ws_doubles(n, K, G > 0 ? G : 1) * sizeof(double);
}
#define DEFINE_RC_DD(SUF, T)                                                                                      \
  int chub_rc_dd_panel_owner_##SUF(void *A, int64_t n, int64_t ld, int G, int rank, int J, void *workspace,       \
                                   double *payload, int32_t *status, void *stream) {                              \
    return rc_dd_panel_owner<T>(A, n, ld, G, rank, J, workspace, payload, status, stream);                        \
  }                                                                                                               \
  int chub_rc_dd_panel_apply_##SUF(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J,         \
                                   void *workspace, const double *payload, int32_t *status, void *stream) {       \
    return rc_dd_panel_apply<T>(A, m_loc, n, ld, G, rank, J, workspace, payload, status, stream);                 \
  }                                                                                                               \
  int chub_rc_dd_finish_##SUF(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v,            \
                              void *workspace, int32_t *status, void *stream) {                                   \
    return rc_dd_finish<T>(A, m_loc, n, ld, G, rank, v, workspace, status, stream);                               \
  }
DEFINE_RC_DD(f64, double)
DEFINE_RC_DD(f32, float)

#define DEFINE_RCW(SUF, T)                                                                                         \
  int chub_rcw_prepare_##SUF(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, const void *V,        \
                             int64_t K, int64_t ldv, void *workspace, int flags, int32_t *status, void *stream) {  \
    return rcw_prepare<T>(A, m_loc, n, ld, G, rank, V, K, ldv, workspace, flags, status, stream);                  \
  }                                                                                                                \
  int chub_rcw_owner_##SUF(void *A, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d, void *workspace, \
                           double *pay_mine, int solve_only, int fuse, int32_t *status, void *stream) {            \
    return rcw_owner<T>(A, n, ld, G, rank, K, d, workspace, pay_mine, nullptr, nullptr, 0, solve_only, fuse,       \
                        status, stream);                                                                           \
  }                                                                                                                \
  int chub_rcw_apply_##SUF(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,   \
                           void *workspace, const double *pay_all, void *V_out, int64_t ldv, int rows_mode,        \
                           void *solve_ws, int fuse, int32_t *status, void *stream) {                              \
    return rcw_apply<T>(A, m_loc, n, ld, G, rank, K, d, workspace, pay_all, nullptr, nullptr, V_out, ldv, 0,       \
                        rows_mode, solve_ws, fuse, status, stream);                                                \
  }                                                                                                                \
  int chub_rcw_owner_p2p_##SUF(void *A, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,              \
                               void *workspace, double *pay_mine, const void *peer_table, void *mailbox,           \
                               int64_t epoch, int solve_only, int fuse, int32_t *status, void *stream) {           \
    if (!peer_table || !mailbox) return (int)cudaErrorInvalidValue;                                                \
    return rcw_owner<T>(A, n, ld, G, rank, K, d, workspace, pay_mine, peer_table, mailbox, epoch, solve_only,      \
                        fuse, status, stream);                                                                     \
  }                                                                                                                \
  int chub_rcw_apply_p2p_##SUF(void *A, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K,          \
                               int64_t d, void *workspace, const void *peer_table, void *mailbox, void *V_out,     \
                               int64_t ldv, int64_t epoch, int rows_mode, void *solve_ws, int fuse,                \
                               int32_t *status, void *stream) {                                                    \
    if (!peer_table || !mailbox) return (int)cudaErrorInvalidValue;                                                \
    return rcw_apply<T>(A, m_loc, n, ld, G, rank, K, d, workspace, (const double *)mailbox, peer_table, mailbox,   \
                        V_out, ldv, epoch, rows_mode, solve_ws, fuse, status, stream);                             \
  }
DEFINE_RCW(f64, double)
DEFINE_RCW(f32, float)

#define DEFINE_SINGLE(NAME, T, UPD)                                                              \
  int NAME(void *L, int64_t n, int64_t ld, int layout, void *v, int flags, void *workspace,      \
           size_t workspace_bytes, int32_t *status, void *stream) {                              \
    return run_single<T, UPD>(L, n, ld, layout, v, flags, workspace, workspace_bytes, status,    \
                              stream);                                                           \
  }
DEFINE_SINGLE(chub_update_f64, double, true)
DEFINE_SINGLE(chub_update_f32, float, true)
DEFINE_SINGLE(chub_downdate_f64, double, false)
DEFINE_SINGLE(chub_downdate_f32, float, false)

#define DEFINE_BATCHED(NAME, T, UPD)                                                             \
  int NAME(void *L, int64_t n, int64_t ld, int64_t stride_L, int layout, void *v,                \
           int64_t stride_v, int64_t batch, int flags, int32_t *status, void *stream) {          \
    return run_batched<T, UPD>(L, n, ld, stride_L, layout, v, stride_v, batch, flags, status,    \
                               stream);                                                          \
  }
DEFINE_BATCHED(chub_update_batched_f64, double, true)
DEFINE_BATCHED(chub_update_batched_f32, float, true)
DEFINE_BATCHED(chub_downdate_batched_f64, double, false)
DEFINE_BATCHED(chub_downdate_batched_f32, float, false)

}  // extern "C"
