/*
 * ltr_b200.h -- C-ABI of the B200-native listwise learning-to-rank loss / NDCG-eval path.
 *
 * Drop-in boundary for the hot path of Haiga/nn-with-pytorch-personalized-losses (pure
 * Python/PyTorch, CPU).  The reference has no FFI today; each entry point below replaces the
 * body of one reference callable (cited as file:line, relative to the reference root) and is
 * what a ctypes / torch-extension binding on the reference side would call (INTEGRATION.md
 * shows the stub).
 *
 * Conventions
 *   - plain pointers + sizes; every data pointer is a DEVICE pointer (fp32 row-major
 *     contiguous unless stated); no torch types, no allocation inside the library;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every call is asynchronous on `stream` and returns 0 on success, a cudaError_t value
 *     or a negative LTR_E* code otherwise; nothing throws across the boundary;
 *   - `workspace` is LTR_WORKSPACE_BYTES of device memory, zero-filled ONCE by the caller
 *     and then reused by every call issued on the same stream (the kernels leave it zeroed);
 *   - `disc_prefix` is the fp64 table P[r] = sum_{q=1..r} 1/log2(1+q), r = 0..disc_len-1
 *     (P[0] = 0), on the device; ltr_fill_disc_prefix() builds it.  It may be NULL, in which
 *     case ideal DCGs always take the sorting path.
 *   - labels equal to `pad_value` mark padded documents (reference: padded_value_indicator).
 */
#ifndef LTR_B200_H_
#define LTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTR_WORKSPACE_BYTES 4096

#define LTR_EINVAL (-1)  /* bad argument (null pointer, n out of range, unknown enum) */
#define LTR_ETOOBIG (-2) /* slate longer than the shared-memory capacity supports   */

/* library / build identification: "ltr_b200 <version> sm_100a" */
const char* ltr_version(void);
/* message for the last non-zero status returned on this thread */
const char* ltr_last_error(void);
/* longest slate the pairwise kernels accept on the current device */
int ltr_max_slate_len(void);

/* P[r] = sum_{q<=r} 1/log2(1+q) in fp64, r = 0..len-1.  Used for O(levels) ideal DCGs. */
int ltr_fill_disc_prefix(double* disc_prefix, int len, void* stream);

/*
 * approxNDCG, fused forward + backward.       replaces losses/approxNDCG.py:7-53
 *   y_pred, y_true : [B, n]
 *   slate_ndcg     : [B]    out, per-slate approximate NDCG (sum_i G_i / log2(1+pos_i))
 *   loss           : [1]    out, -mean_b slate_ndcg[b]               (approxNDCG.py:53)
 *   grad           : [B, n] out, d loss / d y_pred (0 for padded docs); NULL = forward only
 * One CTA per slate, pairs evaluated in registers; no [B,n,n] tensor exists anywhere.
 */
int ltr_approx_ndcg_fwdbwd(const float* y_pred, const float* y_true, int B, int n, float alpha, float eps,
                           float pad_value, const double* disc_prefix, int disc_len, float* slate_ndcg,
                           float* loss, float* grad, void* workspace, void* stream);

/*
 * Same computation from HOST buffers (pinned for asynchrony): H2D of y_pred / y_true, the fused
 * kernel, D2H of loss (+ grad, + per-slate values when the pointers are non-NULL), all queued on
 * `stream`.  d_scratch: ltr_approx_ndcg_host_scratch_bytes(B, n) bytes of device memory.
 */
size_t ltr_approx_ndcg_host_scratch_bytes(int B, int n);
int ltr_approx_ndcg_fwdbwd_host(const float* h_y_pred, const float* h_y_true, int B, int n, float alpha, float eps,
                                float pad_value, const double* disc_prefix, int disc_len, float* h_loss,
                                float* h_grad, float* h_slate_ndcg, void* d_scratch, size_t scratch_bytes,
                                void* workspace, void* stream);

/*
 * Roofline probes (measurement only): a pure ex2.approx loop and a pure FFMA loop, 8 CTAs of
 * 256 threads per SM.  *ops_out (host pointer, may be NULL) receives the number of SFU ops /
 * FP32 FMA lane-ops one launch executes; `out` is one device float that is never written.
 */
int ltr_probe_sfu(float* out, int iters, long long* ops_out, void* stream);
int ltr_probe_fma(float* out, int iters, long long* ops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTR_B200_H_ */
