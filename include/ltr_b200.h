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
 *   - `workspace` is LTR_WORKSPACE_BYTES of 16-byte aligned device memory, zero-filled ONCE by the
 *     caller and then reused by every call issued on the same stream (the kernels leave it zeroed);
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

#define LTR_WORKSPACE_BYTES 262144

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
 *   scratch        : ltr_approx_ndcg_scratch_bytes(B, n) bytes of device memory, or NULL.  With it, batches of few
 *                    long slates (n > 256) spread every slate over ceil(n/128) CTAs (rank launch, gradient launch).
 * Pairs are evaluated in registers; no [B,n,n] tensor exists anywhere.
 */
size_t ltr_approx_ndcg_scratch_bytes(int B, int n);
int ltr_approx_ndcg_fwdbwd(const float* y_pred, const float* y_true, int B, int n, float alpha, float eps,
                           float pad_value, const double* disc_prefix, int disc_len, float* slate_ndcg,
                           float* loss, float* grad, void* scratch, size_t scratch_bytes, void* workspace,
                           void* stream);

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
 * ListNet, fused forward + backward.          replaces losses/listnet.py:5-16   (NB: y_true FIRST, as there)
 *   y_true, y_pred : [B, n], n <= 1024, no padding mask (the reference has none)
 *   slate_loss     : [B] out or NULL, -sum_i softmax(y_true)_i log softmax(y_pred)_i   (apply_sigmoid: -sum sigmoid(.))
 *   loss           : [1] out, sum over the batch                                         (listnet.py:16)
 *   grad           : [B, n] out or NULL, d loss / d y_pred
 * One warp streams one slate; 12 n bytes of HBM traffic per slate.
 */
int ltr_listnet_fwdbwd(const float* y_true, const float* y_pred, int B, int n, int apply_sigmoid, float* slate_loss,
                       float* loss, float* grad, void* workspace, void* stream);

/*
 * Ordinal loss, fused forward + backward.     replaces losses/ordinal.py:9-53
 *   y_pred : [B, n, levels] probabilities, y_true : [B, n]; loss [1]; grad [B, n, levels] or NULL.
 *   A label of -1 makes all targets of the doc -1 (with_ordinals' default); entries whose TARGET equals
 *   pad_indicator are masked (ordinal.py:41).  status : [1] device int out or NULL: bit 0 = a target below 0 reached the
 *   BCE (a -1 label: with_ordinals is called with its default padding value), bit 1 = an element of
 *   y_pred lies outside [0, 1] (or is NaN) -- the two conditions torch's BCELoss raises on; the caller raises.
 */
int ltr_ordinal_fwdbwd(const float* y_pred, const float* y_true, int B, int n, int levels, float pad_indicator,
                       float* loss, int* status, float* grad, void* workspace, void* stream);

/* tie rules of the sorting entry points */
#define LTR_TIES_STABLE 0   /* equal keys in ascending index order: python sorted(reverse=True), metrics.py:51 */
#define LTR_TIES_STDSORT 1  /* torch.sort / argsort on the CPU (libstdc++ std::sort over (key, index)), SURVEY A.3 */
#define LTR_TIES_REVERSED 2 /* equal keys in descending index order: np.argsort(kind="stable")[::-1]              */

/*
 * Descending sort of every row.               replaces Tensor.sort(descending=True) / torch.argsort on [B, n]
 *   perm        : [B, n] int32 out or NULL, perm[b, r] = index of the doc at rank r (NaN first, -inf last)
 *   sorted_keys : [B, n] out or NULL
 */
int ltr_sort_desc(const float* keys, int B, int n, int tie_rule, int* perm, float* sorted_keys, void* stream);

/*
 * NDCG@k, fp32.                               replaces utils/metrics.py:83-104 (torchNdcg)
 *   y_true, y_pred : [Q, n]; ks : nk <= 32 HOST ints (k <= 0 or k > n: no truncation); out : [Q, nk]
 *   disc32 : n device floats, disc32[r] = (float)log2(2 + r) (the reference divides by np.log2(1 + i) rounded to fp32)
 *   gains 2^l - 1, sequential fp32 accumulation, 0/0 = NaN (the 'tensor' return type maps NaN to 0 in the caller).
 *   All k of a sweep are served by one read of the query.  flags: Q + 1 device ints of scratch or NULL (no
 *   initialisation needed; on return flags[1 + q] != 0 marks the queries that took the full sort).  With it, sweeps
 *   whose largest k is <= 10 (n % 4 == 0, 16-byte aligned rows) take the one-thread-per-query path, sweeps up to k = 31
 *   and n <= 1024 the warp-per-query top-k path; neither sorts, and only queries with tied keys among their first
 *   k + 1 ranks (or a NaN / infinite score, or labels that are not small integers) are sorted in full with `tie_rule`:
 *   rows of at most 128 docs inside the same launch (by the warp that finds them), longer rows by a second launch over
 *   the marked queries.
 */
int ltr_ndcg_at_k(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                  const float* disc32, float* out, int* flags, void* stream);

/*
 * The same call with scratch memory for long rows.   replaces utils/metrics.py:83-104 (torchNdcg) at the evaluation shape of
 * BASELINE configs[4] (8192 queries x 1024 docs)
 *   scratch : ltr_ndcg_at_k_scratch_bytes(Q, n) bytes of 16-byte aligned device memory (0 bytes: rows of at most 128 docs, or
 *   not a multiple of 128 -- the call is then ltr_ndcg_at_k), no initialisation needed, or NULL.  With it, rows of 2..16
 *   whole 128-doc chunks (largest k <= 10, flags given) run the label-in-key kernel chunk by chunk -- every chunk a row of its
 *   own, at that kernel's occupancy -- and one thread per query merges the chunks' top-11 lists; queries with (near-)tied keys
 *   among their first k + 1 ranks, odd labels or non-finite scores are marked in `flags` and sorted in full as before.
 */
size_t ltr_ndcg_at_k_scratch_bytes(int Q, int n);
int ltr_ndcg_at_k_ws(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                     const float* disc32, float* out, int* flags, void* scratch, size_t scratch_bytes, void* stream);

/*
 * The same sweep from HOST buffers (pinned for asynchrony): H2D of y_true / y_pred, ltr_ndcg_at_k_ws, D2H of out, all
 * queued on `stream`.  d_scratch: ltr_ndcg_at_k_host_scratch_bytes(Q, n, nk) bytes of 16-byte aligned device memory (the
 * device copies, the flags and the per-chunk lists of long rows); disc32 stays a
 * device table.  This is the call a CPU-resident evaluation loop (main.py:145) makes once per epoch.
 */
size_t ltr_ndcg_at_k_host_scratch_bytes(int Q, int n, int nk);
int ltr_ndcg_at_k_host(const float* h_y_true, const float* h_y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                       const float* disc32, float* h_out, void* d_scratch, size_t scratch_bytes, void* stream);

/*
 * NDCG@k, fp64.                               replaces utils/metrics.py:48-80 (dcg / ndcg / mNdcg)
 *   gains linear (exponential_gains = 0) or 2^l - 1; disc64[r] = log2(2 + r); k clipped to n; a query whose
 *   ideal DCG is 0 gets empty_value (1.0 for no_relevant=True, 0.0 otherwise); out : [Q, nk] doubles.
 */
int ltr_ndcg_at_k_f64(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                      int exponential_gains, double empty_value, const double* disc64, double* out, int* flags,
                      void* stream);

/*
 * DCG@k, fp64 (the un-normalised sum).        replaces utils/metrics.py:48-66 (dcg)
 *   out[q, j] = sum_{r < min(k_j, n)} gain(y_true[q, order_r]) / log2(2 + r), order = y_pred sorted descending with
 *   `tie_rule`; same arguments as ltr_ndcg_at_k_f64 (no empty_value: nothing is divided).
 */
int ltr_dcg_at_k_f64(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                     int exponential_gains, const double* disc64, double* out, int* flags, void* stream);

/*
 * exactNDCG "loss", forward only.             replaces losses/exactNDCG.py:5-29 (ndcgLoss; y_true FIRST)
 *   slate_value [B] = -dcg/idcg on softmaxed labels ordered by softmax(y_pred), discount ln(r) + 1; loss [1] = mean.
 */
int ltr_exact_ndcg(const float* y_true, const float* y_pred, int B, int n, float* slate_value, float* loss, void* stream);

/*
 * orderScore loss, fused forward + backward.  replaces losses/orderScore.py:5-24 (y_true FIRST)
 *   slate_value [B] = -sum_r (-n + 2r) log softmax(y_pred)[order_r], order = argsort(softmax(y_true), descending);
 *   loss [1] = mean; grad [B, n] or NULL.
 */
int ltr_order_score_fwdbwd(const float* y_true, const float* y_pred, int B, int n, float* slate_value, float* loss,
                           float* grad, void* stream);

/* weighing schemes of ltr_lambda_fwdbwd (losses/lambdaL.py:96-127) */
#define LTR_SCHEME_NONE 0        /* weighing_scheme=None, rankNet_scheme                     */
#define LTR_SCHEME_NDCG1 1       /* ndcgLoss1_scheme                                         */
#define LTR_SCHEME_NDCG2 2       /* ndcgLoss2_scheme                                         */
#define LTR_SCHEME_LAMBDARANK 3  /* lamdbaRank_scheme [sic]                                  */
#define LTR_SCHEME_NDCG2PP 4     /* ndcgLoss2PP_scheme                                       */
#define LTR_SCHEME_GTDIFF 5      /* rankNetWeightedByGTDiff_scheme                           */
#define LTR_SCHEME_GTDIFF_POW 6  /* rankNetWeightedByGTDiffPowed_scheme                      */

/*
 * LambdaLoss framework, fused forward + backward.   replaces losses/lambdaL.py:7-93 (lambdaMask + lambdaLoss)
 *   y_pred, y_true : [B, n], n <= 2048; k < 0 (the reference's k=None) or k >= n: no NDCG@k truncation; k = 0: no pair
 *   lies inside the truncation (sum 0, mean NaN), as the reference's [:0, :0] mask gives (lambdaL.py:29-30)
 *   slate_sum, slate_count : [B] out, sum / number of the masked pair log-probabilities of each slate
 *   loss       : [1] out, -sum (reduction_mean = 0) or -sum/count over the whole batch (lambdaL.py:86-89)
 *   pair_count : [1] out or NULL, number of masked pairs in the batch
 *   grad       : [B, n] out or NULL, d loss / d y_pred (0 for padded docs)
 *   workspace  : LTR_WORKSPACE_BYTES of zero-filled device memory per stream (slates are handed out dynamically), or NULL
 */
int ltr_lambda_fwdbwd(const float* y_pred, const float* y_true, int B, int n, float eps, float pad_value, int scheme,
                      int k, float sigma, float mu, int reduction_mean, int natural_log, float* slate_sum,
                      float* slate_count, float* loss, float* pair_count, float* grad, void* workspace, void* stream);

/*
 * Risk-sensitive losses, ListNet family, fused forward + backward.
 * replaces losses/riskLosses/riskLosses.py:8-60 (geoRiskListnetLoss), 128-180 (zRiskListnetLoss), 247-291 (tRiskListnetLoss)
 *
 * Three launches, selected by the bits of `phases`:
 *   1  rows:     mat[b, :] = effectiveness of {model, S baseline systems, (ideal)} for query b --
 *                transformation 1: sum_d (p_t p_s - p_t^2)^2   2: cosine(p_t, p_s)   3: (sum p_t p_s - sum p_t^2)^2,
 *                p = softmax over the docs; kind 2 (t-risk) uses cosine(p_t^2, p_t p_s) for transformation 2 (:262) --
 *                and this rank's phase-1 state (matrix max, column totals, query count) -> exchange block, round 0;
 *   2  stats:    (needs round 0 of every rank) z-risk sums -> exchange block, round 1; per-query gradient terms -> rowstate;
 *   4  backward: (needs both rounds of every rank) loss[0]; grad[b, :] = d loss / d y_pred[b, :] (grad may be NULL).
 * kind 0 zRisk, 1 geoRisk, 2 tRisk (columns 0 = model, 1 = baseline); invert: mat <- max(mat) - mat (:49-50);
 * return_strategy 1: f r(0), 2: f (r(-1) - r(0)), 3: f (r(-1) - r(0))^2; precedence_quirk: strategy 2 is
 * f r(-1) - r(0) (zRiskListnetLoss, :176); negative = f.
 *   y_baselines : [B, n, S] or NULL (S = 0); mat : [B, 1 + S + add_ideal] out; rowstate : 3 B device doubles of scratch.
 *   B = 0 is an empty shard: the rank contributes the neutral state (its data pointers may be NULL).
 *
 * Query sharding.  xchg: this rank's exchange block, LTR_RISK_XCHG_BYTES of device memory zero-filled
 * once (one per stream): inbox[round][rank][16] doubles + flags.  world = 1: phases = 7, one call.  world > 1, either
 *   - peer_xchg = NULL: call with phases = 1, all-gather row `rank` of round 0 (16 doubles at byte offset
 *     (0 * LTR_RISK_MAX_WORLD + rank) * 128) into everybody's block, phases = 2, all-gather round 1 (offset
 *     (1 * LTR_RISK_MAX_WORLD + rank) * 128), phases = 4: two collectives per step; or
 *   - peer_xchg = HOST array of `world` device pointers, every rank's block as mapped into this process
 *     (ltr_risk_xchg_alloc + ltr_ipc_export / ltr_ipc_open; peer_xchg[rank] == xchg): phases = 7, one call, no
 *     collective -- the kernels store their states into the peers' inboxes over NVLink and spin (<= 4 s) on flags
 *     carrying the call's sequence number, which the kernels keep in the exchange block themselves (every rank must
 *     make the same sequence of calls on its block; the three launches are CUDA-graph capturable).
 */
#define LTR_RISK_MAX_WORLD 16
#define LTR_RISK_XCHG_BYTES 16384
int ltr_risk_listnet_fwdbwd(const float* y_pred, const float* y_true, const float* y_baselines, int B, int n, int S,
                            int transformation, int add_ideal, int kind, int invert, int return_strategy,
                            int precedence_quirk, float alpha, float negative, float* mat, double* rowstate, float* loss,
                            float* grad, int phases, void* xchg, int rank, int world, const void* const* peer_xchg,
                            void* workspace, void* stream);

/*
 * Exchange blocks in peer-mappable memory.  ltr_risk_xchg_alloc: cudaMalloc + zero fill of LTR_RISK_XCHG_BYTES on the
 * current device (the one allocation this library makes: CUDA IPC needs the base of an allocation);
 * ltr_ipc_export: the 64-byte cudaIpcMemHandle_t of a block; ltr_ipc_open: map a peer's block (another process, same
 * or another GPU of the box; enables peer access); ltr_risk_xchg_status: 0, or 1 when a peer's state did not arrive
 * within the spin limit (the loss of that call is NaN) -- synchronous, for diagnostics.
 */
int ltr_risk_xchg_alloc(void** xchg);
int ltr_risk_xchg_free(void* xchg);
int ltr_ipc_export(void* dptr, unsigned char* handle64);
int ltr_ipc_open(const unsigned char* handle64, void** dptr);
int ltr_ipc_close(void* dptr);
int ltr_risk_xchg_status(const void* xchg, int* status);

/*
 * Risk-sensitive losses, LambdaLoss family.   replaces losses/riskLosses/riskLosses.py:63-125, 183-244, 294-345
 *
 * ltr_risk_lambda_rows: per system, the column sums (over dim 1) of lambdaMask(softmax(scores), softmax(y_true),
 *   weighing_scheme, return_losses=True) -- the UNMASKED pair log-probabilities, lambdaMask's default eps / sigma /
 *   mu -- are compared with those of the labels ranked by themselves:
 *   transformation 1: sum_j (c_j - base_j)^2   2: cosine(base, c)   3 (t-risk only): (sum c - sum base)^2
 *   mat : [B, 1 + S + add_ideal]; ideal_ones: the ideal column of transformation 2 is 1 (geoRisk, :107) instead of
 *   cosine(base, base) (zRisk, :227); dvec : [B, n] out or NULL, d mat[b, 0] / d (model column sum j), rank order.
 * ltr_risk_lambda_rows_bwd: grad[b, :] = gq[b] * d mat[b, 0] / d y_pred[b, :] (pair pass + softmax Jacobian).
 */
int ltr_risk_lambda_rows(const float* y_pred, const float* y_true, const float* y_baselines, int B, int n, int S,
                         int transformation, int add_ideal, int ideal_ones, int scheme, float* mat, float* dvec,
                         void* stream);
int ltr_risk_lambda_rows_bwd(const float* y_pred, const float* y_true, int B, int n, int scheme, const float* gq,
                             const float* dvec, float* grad, void* stream);

/*
 * lambdaMask.                                 replaces losses/lambdaL.py:7-64
 *   logp : [B, n, n] out, log-probability of every pair of rank positions (the return_losses=True result);
 *   mask : [B, n, n] bytes out or NULL, 1 where the pair enters lambdaLoss (label_i > label_j, both real, both < k).
 *   k as in ltr_lambda_fwdbwd (k < 0: none).
 * ltr_lambda_pair_matrix_bwd: grad [B, n] = d L / d y_pred given grad_logp [B, n, n] = d L / d logp (rank order); the
 *   weights carry no gradient (lambdaL.py:44-47), padded docs get 0.
 */
int ltr_lambda_pair_matrix(const float* y_pred, const float* y_true, int B, int n, float eps, float pad_value, int scheme,
                           int k, float sigma, float mu, int natural_log, float* logp, unsigned char* mask,
                           void* stream);
int ltr_lambda_pair_matrix_bwd(const float* y_pred, const float* y_true, int B, int n, float eps, float pad_value,
                               int scheme, int k, float sigma, float mu, int natural_log, const float* grad_logp,
                               float* grad, void* stream);

/*
 * zRisk / geoRisk / tRisk over a caller-provided query x system matrix.   replaces losses/riskLosses/riskFunctions.py:4-33
 * and the epilogues riskLosses.py:49-58, 274-291 (used by zRisk / geoRisk themselves, the LambdaLoss-style risk losses and
 * the validation epilogue).  The same three phases as above, one single-CTA launch each, states in the exchange block:
 *   phase 1 (local)  round 0 <- {max(mat), sum mat[:,0], sum mat[:,C-1], sum(mat), Q}          (combined: MAX on [0], SUM)
 *   phase 2          round 1 <- per used column {sum z c, sum A T, sum A e, sum c/sqrt(e), sum A}, #maxima, t-risk sums
 *   phase 3          loss[0], gq[q] = d loss / d mat[q, 0] and gmat[q, i] = d loss / d mat[q, i]   (either may be NULL)
 * Query-sharded callers all-gather the inbox rows between the phases as described above; Q = 0 is an empty shard.
 */
int ltr_risk_phase(int phase, const float* mat, int Q, int C, int kind, int invert, int return_strategy,
                   int precedence_quirk, float alpha, float negative, void* xchg, int rank, int world, float* loss,
                   float* gq, float* gmat, void* stream);

/*
 * GeoRisk of every system column, fp64.       replaces utils/metrics.py:8-45 (getGeoRiskDefault)
 *   mat : [Q, C] device DOUBLES (queries x systems); out : [C] device doubles.  An entry whose expected value
 *   S_i T_q / N is 0 contributes z = 0 (the reference's guard, :31-34); z is scaled by 1 + alpha where x < e.
 */
int ltr_georisk_default_f64(const double* mat, int Q, int C, double alpha, double* out, void* stream);

/*
 * Fused ranker tail + loss: TripleLayerNet forward, approxNDCG forward + backward and the ranker's weight
 * gradients in one pass over the features.   replaces, for BASELINE configs[1], the step
 *   architeture/tripleLayer.py:5-17 (forward) -> losses/approxNDCG.py:7-53 -> loss.backward()  (main_batch_execution.py:128-170)
 * l1 (no activation) and l2 are folded by the caller: Wc = W2 W1 [32, F], bc = W2 b1 + b2; z = x Wc^T + bc, a = sigmoid(z),
 * score = a . w3 + b3.  tcgen05 tf32 tensor-core MMAs (fp32 accumulation in tensor memory), feature tiles by TMA.
 *   X : [B, 128, F] features (F <= 160, F % 4 == 0, 16-byte aligned), y_true : [B, 128] (pad_value marks padding)
 *   scores [B, 128], slate_ndcg [B], loss [1] out;  gWc [32, F], gbc [32], gw3 [32], gb3 [1] out:
 *   d loss / d {Wc, bc, w3, b3}; the caller unfolds them into the six parameter gradients
 *   (dW1 = W2^T gWc, db1 = W2^T gbc, dW2 = gWc W1^T + gbc b1^T, db2 = gbc, dW3 = gw3, db3 = gb3).
 *   scratch : ltr_triple_scratch_bytes(B, 128, F) bytes of device memory (per-CTA partial gradients).
 */
size_t ltr_triple_scratch_bytes(int B, int n, int F);
int ltr_triple_approx_ndcg_fwdbwd(const float* X, const float* y_true, const float* Wc, const float* bc, const float* w3,
                                  float b3, int B, int n, int F, float alpha, float eps, float pad_value,
                                  const double* disc_prefix, int disc_len, float* scores, float* slate_ndcg, float* loss,
                                  float* gWc, float* gbc, float* gw3, float* gb3, void* scratch, size_t scratch_bytes,
                                  void* workspace, void* stream);

/*
 * The same step from the module's own six parameters, in one call: the fold (Wc = W2 W1, bc = W2 b1 + b2), the fused
 * kernel and the unfold of the gradients all run on the device (four launches chained by programmatic stream
 * serialization); nothing is read back to the host.      architeture/tripleLayer.py:8-10 (l1: F -> 64, l2: 64 -> 32, l3: 32 -> 1)
 *   W1 [64, F], b1 [64], W2 [32, 64], b2 [32], W3 [1, 32], b3 [1] device pointers; gW1 .. gb3 the matching gradients (out).
 *   scratch : ltr_triple_step_scratch_bytes(B, 128, F) bytes, 16-byte aligned.
 */
size_t ltr_triple_step_scratch_bytes(int B, int n, int F);
int ltr_triple_approx_ndcg_step(const float* X, const float* y_true, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* W3, const float* b3, int B, int n, int F, float alpha, float eps,
                                float pad_value, const double* disc_prefix, int disc_len, float* scores, float* slate_ndcg,
                                float* loss, float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3,
                                void* scratch, size_t scratch_bytes, void* workspace, void* stream);

/*
 * The same fused step with the ListNet loss instead of approxNDCG.   replaces, for a TripleLayerNet ranker,
 *   architeture/tripleLayer.py:5-17 -> losses/listnet.py:5-16 (apply_sigmoid = False) -> loss.backward()
 * loss = - sum over the batch of sum_i softmax(y_true)_i log softmax(score)_i over all 128 docs of a slate (the reference gives
 * padding no special treatment); slate_loss [B]: the per-slate sums.  Arguments as ltr_triple_approx_ndcg_step.
 */
int ltr_triple_listnet_step(const float* X, const float* y_true, const float* W1, const float* b1, const float* W2,
                            const float* b2, const float* W3, const float* b3, int B, int n, int F, const double* disc_prefix,
                            int disc_len, float* scores, float* slate_loss, float* loss, float* gW1, float* gb1, float* gW2,
                            float* gb2, float* gW3, float* gb3, void* scratch, size_t scratch_bytes, void* workspace,
                            void* stream);

/*
 * Average of small tensors over the GPUs of one box through peer memory (no collective library call): the gradient
 * exchange that follows the fused ranker step under data parallelism.   replaces DistributedDataParallel's gradient
 * all-reduce in the reference's loop (main_batch_execution.py:120-171: ~13 k floats per step for TripleLayerNet)
 *   ptrs / counts : nseg <= 8 device segments, at most ltr_peer_xchg_max_floats() floats in total; averaged IN PLACE,
 *                   every rank gets the same bits.   xchg : this rank's block (ltr_peer_xchg_alloc: ltr_peer_xchg_bytes()
 *   of cudaMalloc'd, zero-filled memory); peer_xchg[world] : every rank's block as mapped into this process (ltr_ipc_export /
 *   ltr_ipc_open; peer_xchg[rank] == xchg).  ONE launch on `stream` (CTA b pushes values [1024 b, 1024 b + 1024) to every peer, then gathers them), capturable in a CUDA graph; every rank
 *   must make the same sequence of calls.  ltr_peer_xchg_status: 1 after a peer failed to arrive within 4 s (results NaN).
 */
size_t ltr_peer_xchg_bytes(void);
int ltr_peer_xchg_max_floats(void);
int ltr_peer_xchg_alloc(void** xchg);
int ltr_peer_xchg_free(void* xchg);
int ltr_peer_xchg_status(const void* xchg, int* status);
int ltr_peer_allreduce_avg(float* const* ptrs, const int* counts, int nseg, void* xchg, void* const* peer_xchg, int rank,
                           int world, void* stream);

/*
 * Slate ingestion (HOST code, host pointers).  replaces utils/dataset.py:33-101 (get_data, load_baseline_file)
 *
 * svmlight / LETOR text "label qid:ID idx:val ... # comment" -> query-major padded arrays.  A new query starts wherever
 * the qid of consecutive lines CHANGES (as the reference groups them, dataset.py:52-65).
 *   ltr_svmlight_scan: counts[5] = {lines, queries, docs in the longest query, smallest feature index, largest}.
 *   ltr_svmlight_load: queries [first_query, first_query + num_queries) -- a rank's shard -- into
 *     X [num_queries, n_docs, n_features] (dense, zero-filled), y [num_queries, n_docs] (pad_value beyond a query's
 *     docs), qids [num_queries] and lens [num_queries] (docs of the query before truncation to n_docs; either may be
 *     NULL).  index_base 1 (LETOR), 0, or -1 = sklearn's zero_based="auto" (zero-based iff an index 0 occurs).
 *     `threads` host threads parse the lines (0 = all cores).  Buffers are the caller's (pinned for an async H2D copy).
 * Baseline scores "qid,score_1,...,score_S" per doc:
 *   ltr_baseline_csv_scan: counts[4] = {lines, queries, docs in the longest query, S}.
 *   ltr_baseline_csv_load: scores [num_queries, n_docs, S] (zero beyond a query's docs), lens [num_queries] or NULL.
 */
int ltr_svmlight_scan(const char* path, int threads, long long* counts);
int ltr_svmlight_load(const char* path, long long first_query, long long num_queries, int n_docs, int n_features,
                      int index_base, float pad_value, float* X, float* y, long long* qids, int* lens, int threads);
int ltr_baseline_csv_scan(const char* path, long long* counts);
int ltr_baseline_csv_load(const char* path, long long first_query, long long num_queries, int n_docs, int S, float* scores,
                          int* lens);

/*
 * Roofline probes (measurement only): a pure ex2.approx loop and a pure FFMA loop, 8 CTAs of
 * 256 threads per SM.  *ops_out (host pointer, may be NULL) receives the number of SFU ops /
 * FP32 FMA lane-ops one launch executes; `out` is one device float that is never written.
 */
int ltr_probe_sfu(float* out, int iters, long long* ops_out, void* stream);
int ltr_probe_fma(float* out, int iters, long long* ops_out, void* stream);
/* the feature-tile TMA stream of the fused ranker step on its own: every CTA keeps `slots` (<= 13) 16 KB copies in flight
 * over its slates of X [B, 128, F], `reps` times over (a small B stays in L2); mode 0 = [128 docs][32 floats] tensor-map
 * boxes, 1 = contiguous 4 KB bulk copies, 2 = [32 docs][all chunks] slots of 4 KB boxes with 32-byte swizzle atoms (slots <= 10) */
int ltr_probe_tma_stream(const float* X, int B, int F, int slots, int mode, int reps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTR_B200_H_ */
