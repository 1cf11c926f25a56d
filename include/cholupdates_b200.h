/*
 * cholupdates_b200.h -- C ABI of the B200 (sm_100a) rank-1 Seeger Cholesky
 * update / downdate.
 *
 * This is the drop-in boundary for ONE hot path of marvinpfoertner/cholupdates:
 * the inner kernels
 *
 *     cpdef void update  (blas_dtype[:, :] L, blas_dtype[::1] v) except *
 *         src/cholupdates/rank_1/_seeger_impl_cython.pyx:22-155
 *     cpdef void downdate(blas_dtype[:, :] L, blas_dtype[::1] v) except *
 *         src/cholupdates/rank_1/_seeger_impl_cython.pyx:160-330
 *
 * (fused type blas_dtype in {float, double}, src/cholupdates/_blas_polymorphic.pxd:8-12)
 * which the reference front-end selects as `_update_impl_default` /
 * `_downdate_impl_default` (src/cholupdates/rank_1/_seeger.py:17-34) and calls at
 * _seeger.py:279-294 / :715-730.  Like those kernels, the entry points below
 * mutate L and v in place and do NO argument validation (shape / dtype /
 * contiguity checks stay in the host front-end, exactly as in the reference,
 * _seeger.py:241-269 + _arg_validation.py:6-30).
 *
 * Conventions
 * -----------
 *  - Plain C: pointers and sizes only, no CUDA or torch types.  `stream` is a
 *    cudaStream_t passed as void* (NULL = the legacy default stream).
 *  - All functions return 0 on success or a cudaError_t-style non-zero code
 *    (see chub_last_error()).  Numerical outcomes are reported through
 *    `status`, not through the return value.
 *  - L is an n x n matrix with leading dimension `ld` (elements), element
 *    (i, j) at L[i*ld + j] (CHUB_ROW_MAJOR, NumPy "C") or L[i + j*ld]
 *    (CHUB_COL_MAJOR, NumPy "F").  Only i >= j is ever read or written
 *    (reference contract, tests/cholupdates/rank_1/update/test_update.py:25-55).
 *  - v has n contiguous elements.  On return it holds scratch: v[k] = the BLAS
 *    rotg reconstruction parameter z_k, which is what the reference's Cython
 *    path leaves there (.pyx:101,152 / :288).  After a CHUB_STATUS_NOT_PD
 *    downdate it holds p = L^{-1} v (.pyx:239-259) and L is untouched.
 *  - status codes (int32, written by the device):
 *        CHUB_STATUS_OK         0
 *        CHUB_STATUS_ZERO_DIAG  1   `check_diag` found L[k][k] == 0
 *                                   (_arg_validation.py:18-22); L and v untouched
 *        CHUB_STATUS_NOT_PD     2   downdate: rho^2 = 1 - p.p <= 0 (.pyx:255-259);
 *                                   L untouched
 *        CHUB_STATUS_INTERNAL   3   a device-side watchdog fired (never expected)
 */
#ifndef CHOLUPDATES_B200_H
#define CHOLUPDATES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHUB_ROW_MAJOR 0
#define CHUB_COL_MAJOR 1

#define CHUB_FLAG_CHECK_DIAG 1 /* reference kwarg check_diag (_seeger.py:40) */

#define CHUB_STATUS_OK 0
#define CHUB_STATUS_ZERO_DIAG 1
#define CHUB_STATUS_NOT_PD 2
#define CHUB_STATUS_INTERNAL 3

#define CHUB_DTYPE_F32 0
#define CHUB_DTYPE_F64 1

/* Library / device information. */
int chub_version(void);
const char *chub_last_error(void);
int chub_device_count(void);
/* SM count and L2 size of the current device (used by bench.py for grid / flush sizing). */
int chub_device_info(int *sm_count, int64_t *l2_bytes, int *cc_major, int *cc_minor);

/* Bytes of device workspace the device-pointer entry points need for a factor of
 * order n (per call, independent of batch).  Pass a buffer of at least this size,
 * or workspace == NULL to let the library use an internal per-device buffer.
 * The internal buffer is ONE per device: calls that rely on it must be issued on a
 * single stream (or serialised by the caller); concurrent streams / threads each pass
 * their own workspace (the Python front-end keeps one per (device, stream)). */
size_t chub_workspace_bytes(int dtype, int64_t n);

/* ---- device-pointer entry points: L, v, workspace, status are DEVICE pointers;
 *      asynchronous on `stream`.  `status` points to one int32.
 *      Replace _seeger_impl_cython.update / .downdate (.pyx:22-155, :160-330). */
int chub_update_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                    void *workspace, size_t workspace_bytes, int32_t *status, void *stream);
int chub_update_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                    void *workspace, size_t workspace_bytes, int32_t *status, void *stream);
int chub_downdate_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                      void *workspace, size_t workspace_bytes, int32_t *status, void *stream);
int chub_downdate_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                      void *workspace, size_t workspace_bytes, int32_t *status, void *stream);

/* ---- batched entry points: `batch` independent factors, factor b at
 *      L + b*stride_L, v + b*stride_v (elements); status[batch], one code per
 *      factor (a failing factor is left untouched, the others proceed).
 *      New surface (the reference has no batched call; a 3-D L must keep raising
 *      ValueError through update(), test_update.py:88-97). */
int chub_update_batched_f64(void *L, int64_t n, int64_t ld, int64_t stride_L, int layout, void *v,
                            int64_t stride_v, int64_t batch, int flags, int32_t *status,
                            void *stream);
int chub_update_batched_f32(void *L, int64_t n, int64_t ld, int64_t stride_L, int layout, void *v,
                            int64_t stride_v, int64_t batch, int flags, int32_t *status,
                            void *stream);
int chub_downdate_batched_f64(void *L, int64_t n, int64_t ld, int64_t stride_L, int layout, void *v,
                              int64_t stride_v, int64_t batch, int flags, int32_t *status,
                              void *stream);
int chub_downdate_batched_f32(void *L, int64_t n, int64_t ld, int64_t stride_L, int layout, void *v,
                              int64_t stride_v, int64_t batch, int flags, int32_t *status,
                              void *stream);

/* ---- host-buffer entry points: L and v are HOST pointers (what the reference's
 *      NumPy callers hold, _seeger.py:37-44).  Synchronous: the lower triangle is
 *      staged to the current device in row blocks, updated and copied back (the
 *      strict upper triangle of the host array is never read or written);
 *      *status is a host int32.  The downdate copies L back only if status == 0. */
int chub_update_host_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                         int32_t *status);
int chub_update_host_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                         int32_t *status);
int chub_downdate_host_f64(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                           int32_t *status);
int chub_downdate_host_f32(void *L, int64_t n, int64_t ld, int layout, void *v, int flags,
                           int32_t *status);
/* Copy mode of the host entry points (reference: overwrite_L=False copies L first, _seeger.py:272-273 /
 * :708-709): L_in is only read; the updated lower triangle goes to L_out (same layout, own leading
 * dimension), and the strict upper triangle of L_in is copied to L_out on the host, overlapped with the
 * transfers -- no host-side copy of the lower triangle, nothing above the diagonal crosses PCIe.
 * L_out's contents are unspecified when *status != CHUB_STATUS_OK. */
int chub_update_host_copy_f64(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                              void *v, int flags, int32_t *status);
int chub_update_host_copy_f32(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                              void *v, int flags, int32_t *status);
int chub_downdate_host_copy_f64(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                                void *v, int flags, int32_t *status);
int chub_downdate_host_copy_f32(const void *L_in, int64_t ld_in, void *L_out, int64_t ld_out, int64_t n, int layout,
                                void *v, int flags, int32_t *status);
/* Release the staging buffers the host entry points cache on the current device. */
int chub_host_release(void);

/* ---- block-row-cyclic distribution of ONE factor over G devices (row-major).
 *      New surface: the reference has no distributed mode.  Rank g stores the global
 *      256-row blocks g, g+G, g+2G, ... back to back as A_loc[m_loc][ld] (n columns).  Per
 *      256-column panel J the owner of the diagonal block (rank J mod G) forms the panel's
 *      coefficients and packs them into `payload` (chub_rc_payload_doubles() doubles); the caller
 *      broadcasts the payload (NCCL) and every rank applies the panel to its local rows.
 *      `v` / `v_out` are full n-vectors on the device; workspace as for the single-device calls. */
int chub_rc_payload_doubles(void);
int chub_rc_block_rows(void);
int chub_rc_prepare_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v,
                        void *workspace, int flags, int32_t *status, void *stream);
int chub_rc_prepare_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v,
                        void *workspace, int flags, int32_t *status, void *stream);
int chub_rc_panel_owner_f64(void *A_loc, int64_t n, int64_t ld, int G, int rank, int J, void *v_out,
                            void *workspace, double *payload, int32_t *status, void *stream);
int chub_rc_panel_owner_f32(void *A_loc, int64_t n, int64_t ld, int G, int rank, int J, void *v_out,
                            void *workspace, double *payload, int32_t *status, void *stream);
int chub_rc_panel_apply_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J,
                            int is_owner, void *workspace, const double *payload, int32_t *status,
                            void *stream);
int chub_rc_panel_apply_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J,
                            int is_owner, void *workspace, const double *payload, int32_t *status,
                            void *stream);

/* ---- block-row-cyclic DOWNDATE (.pyx:160-330 distributed).  After chub_rc_prepare: per
 *      256-column panel J the owner solves its diagonal block for the panel's p = L^{-1} v
 *      (chub_rc_dd_panel_owner_* -> payload: p and the signs of the input diagonal), the caller
 *      broadcasts the payload, and every rank subtracts L[rows, panel] p from the residual of its rows
 *      below the block (chub_rc_dd_panel_apply_*; read-only on A).  chub_rc_dd_finish then forms
 *      rho^2 = 1 - p.p on every rank (identical arithmetic, no exchange): if rho^2 <= 0 the status
 *      becomes CHUB_STATUS_NOT_PD, v <- p and A stays untouched (.pyx:255-259); otherwise
 *      v <- rotg's z_k and the reverse sweep (.pyx:279-330) runs on the local rows. */
int chub_rc_dd_panel_owner_f64(void *A_loc, int64_t n, int64_t ld, int G, int rank, int J, void *workspace,
                               double *payload, int32_t *status, void *stream);
int chub_rc_dd_panel_owner_f32(void *A_loc, int64_t n, int64_t ld, int G, int rank, int J, void *workspace,
                               double *payload, int32_t *status, void *stream);
int chub_rc_dd_panel_apply_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J,
                               void *workspace, const double *payload, int32_t *status, void *stream);
int chub_rc_dd_panel_apply_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int J,
                               void *workspace, const double *payload, int32_t *status, void *stream);
int chub_rc_dd_finish_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v,
                          void *workspace, int32_t *status, void *stream);
int chub_rc_dd_finish_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, void *v,
                          void *workspace, int32_t *status, void *stream);

/* ---- K successive updates of one block-row-cyclic factor as a wavefront (BASELINE.json
 *      configs[4]).  Task (u, J) = panel J of update u depends only on (u-1, J) and
 *      (u, J-1), so the tasks of an anti-diagonal d = u + J are independent.  Per step d the caller
 *      runs chub_rcw_owner_* for this rank's tasks (-> pay_mine: chub_rcw_max_tasks() slots of
 *      chub_rcw_payload_doubles() doubles), ONE all-gather of pay_mine into pay_all
 *      (G * max_tasks slots, rank-major), then chub_rcw_apply_* for all tasks of the step on the local
 *      rows: K + ceil(n/256) - 1 exchanges instead of K * ceil(n/256).  V / V_out: K vectors of n
 *      elements, leading dimension ldv (V_out may be NULL; it receives rotg's z_k, as v does in the
 *      reference, .pyx:101,152).  With G == 1 pay_all == pay_mine and no exchange is needed.
 *      rows_mode of the apply calls: 0 = all local rows of every task; look-ahead split: 1 = only the
 *      two 256-row blocks J, J+1 of every task (all that the owner call of step d+1 depends on),
 *      2 = the blocks below.  Mode 2 of step d may run on a second stream concurrently with the owner
 *      and mode-1 calls of step d+1; the mode-1 call of step d+1 must wait for mode 2 of step d.
 *      Forward substitution only (the distributed downdate's p = L^{-1} v, K = 1): owner calls with
 *      solve_only = 1, apply calls with solve_ws = the chub_workspace_bytes() workspace that
 *      chub_rc_dd_finish will read (p and the signs of the input diagonal are stored there); A_loc
 *      is not modified.  solve_ws = NULL / solve_only = 0: the update.
 *      fuse = 1: one update per task.  fuse = 2: a task is panel J of a GROUP of two consecutive
 *      updates (2u, 2u+1) -- the panel is read and written once for both, so the trail moves half the
 *      bytes per update; the steps then run d = 0 .. ceil(K/2) + ceil(n/256) - 2.  The owner applies
 *      the first update of the group to a scratch copy of its diagonal block to get the second one's
 *      coefficients; the trail applies both to every row. */
int chub_rcw_payload_doubles(void); /* doubles per task slot (room for chub_rcw_max_fuse() updates) */
int chub_rcw_max_fuse(void);
int chub_rcw_max_tasks(int64_t n, int64_t K, int G);
size_t chub_rcw_workspace_bytes(int64_t n, int64_t K, int G);
int chub_rcw_prepare_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, const void *V,
                         int64_t K, int64_t ldv, void *workspace, int flags, int32_t *status, void *stream);
int chub_rcw_prepare_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, const void *V,
                         int64_t K, int64_t ldv, void *workspace, int flags, int32_t *status, void *stream);
int chub_rcw_owner_f64(void *A_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
                       void *workspace, double *pay_mine, int solve_only, int fuse, int32_t *status,
                       void *stream);
int chub_rcw_owner_f32(void *A_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
                       void *workspace, double *pay_mine, int solve_only, int fuse, int32_t *status,
                       void *stream);
int chub_rcw_apply_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
                       void *workspace, const double *pay_all, void *V_out, int64_t ldv, int rows_mode,
                       void *solve_ws, int fuse, int32_t *status, void *stream);
int chub_rcw_apply_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
                       void *workspace, const double *pay_all, void *V_out, int64_t ldv, int rows_mode,
                       void *solve_ws, int fuse, int32_t *status, void *stream);

/* ---- the same wavefront with the exchange over NVLink peer memory instead of a collective
 *      (one process per GPU).  Every rank allocates a mailbox (chub_p2p_alloc of
 *      chub_rcw_mailbox_bytes), exports its cudaIpc handle (64 bytes), imports the peers' handles
 *      and passes the G mapped base pointers (own mailbox at index `rank`) as a DEVICE array
 *      `peer_table`.  Per step the owner kernel stores its payloads straight into every peer's mailbox
 *      and raises a flag there; the trail kernel waits for the flags in its own mailbox: 2 launches
 *      per step, no collective, no host synchronisation.  `epoch` = number of steps of all earlier
 *      calls on this mailbox (identical on every rank); every rank must call owner and apply for
 *      every step d = 0 .. K + ceil(n/256) - 2.  A peer that never arrives trips a device watchdog
 *      (CHUB_STATUS_INTERNAL) instead of hanging. */
size_t chub_rcw_mailbox_bytes(int64_t n, int64_t K, int G);
int chub_p2p_alloc(size_t bytes, void **ptr);
int chub_p2p_free(void *ptr);
int chub_p2p_export(void *ptr, void *handle64);
int chub_p2p_import(const void *handle64, void **peer_ptr);
int chub_p2p_close(void *peer_ptr);
int chub_rcw_owner_p2p_f64(void *A_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
                           void *workspace, double *pay_mine, const void *peer_table, void *mailbox,
                           int64_t epoch, int solve_only, int fuse, int32_t *status, void *stream);
int chub_rcw_owner_p2p_f32(void *A_loc, int64_t n, int64_t ld, int G, int rank, int64_t K, int64_t d,
                           void *workspace, double *pay_mine, const void *peer_table, void *mailbox,
                           int64_t epoch, int solve_only, int fuse, int32_t *status, void *stream);
int chub_rcw_apply_p2p_f64(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K,
                           int64_t d, void *workspace, const void *peer_table, void *mailbox, void *V_out,
                           int64_t ldv, int64_t epoch, int rows_mode, void *solve_ws, int fuse,
                           int32_t *status, void *stream);
int chub_rcw_apply_p2p_f32(void *A_loc, int64_t m_loc, int64_t n, int64_t ld, int G, int rank, int64_t K,
                           int64_t d, void *workspace, const void *peer_table, void *mailbox, void *V_out,
                           int64_t ldv, int64_t epoch, int rows_mode, void *solve_ws, int fuse,
                           int32_t *status, void *stream);

/* ---- tuning / introspection (bench.py and tests) */
/* Force a kernel path: 0 = automatic, 1 = single-CTA kernels only, 2 = panel
 * (multi-kernel) path, 3 = persistent dataflow kernel.  Returns the previous value. */
int chub_set_path(int path);
/* Debug: per-CTA cycle counters of the persistent kernel (enabled by env CHUB_PROFILE=1),
 * out[cta * 16 + slot]; workspace == NULL reads the internal workspace. */
int chub_debug_profile(void *workspace, int64_t n, long long *out, int ctas);
/* Number of kernels launched by this library since the counter was last reset. */
int64_t chub_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* CHOLUPDATES_B200_H */
