// K2b: the LambdaLoss-style risk losses (replaces losses/riskLosses/riskLosses.py:63-125, 183-244, 294-345) and
// lambdaMask itself (losses/lambdaL.py:7-64).
//
// The reference builds, per system, the UNMASKED [B, n, n] tensor of pair log-probabilities with
// `lambdaMask(softmax(scores), softmax(y_true), return_losses=True)` and sums it over dim 1; a query's
// effectiveness is (column sums of the system - column sums of the labels ranked by themselves)^2 summed over
// the docs (transformation 1), their cosine (2) or the squared difference of their totals (3, t-risk only).
// Here one CTA per query does all systems in turn -- softmax, sort with torch.sort's tie order, gains, then
// thread j owns column j and walks the rows -- and writes one row of the query x system matrix plus
// dvec[b, j] = d mat[b, 0] / d (model column sum j) for the backward pass.  The backward pass walks the
// pairs once more with the upstream weights and finishes with the softmax Jacobian.
//   ltr_lambda_pair_matrix dumps the pair matrix and the mask for lambdaMask's two return modes.
#include "lambda_common.cuh"
#include "ltr_host.h"
#include "sort_common.cuh"

namespace ltr {

constexpr int LR_THREADS = 128;
constexpr int LR_MAX_N = 1024;

struct LambdaRiskParams {
    const float* y_pred;   // [B, n]
    const float* y_true;   // [B, n]
    const float* y_base;   // [B, n, S] or null
    int B, n, len, S, lt, add_ideal, ideal_ones;
    float eps, sigma, mu;
    float* mat;            // [B, C]
    float* dvec;           // [B, n]
    const float* gq;       // backward: [B]
    float* grad;           // backward: [B, n]
};

struct LrSmem {
    float *K, *K0, *G, *Y, *invD, *dtab, *pt, *base, *cur;
    int* I;
    __device__ LrSmem(void* raw, int len) {
        K = reinterpret_cast<float*>(raw);
        I = reinterpret_cast<int*>(K + len);
        K0 = reinterpret_cast<float*>(I + len);
        G = K0 + len;
        Y = G + len;
        invD = Y + len;
        dtab = invD + len;
        pt = dtab + len;
        base = pt + len;
        cur = base + len;
    }
};
constexpr int LR_ARRAYS = 10;

__device__ __forceinline__ float lr_block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}
__device__ __forceinline__ float lr_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
}

// dst[i] = softmax(src[i * stride]) over i < n; returns nothing (block-wide, ends with a barrier)
__device__ __forceinline__ void lr_softmax(const float* src, int stride, int n, float* dst, float* red) {
    const float ninf = __int_as_float(0xff800000);
    float m = ninf;
    for (int i = threadIdx.x; i < n; i += LR_THREADS) m = fmaxf(m, src[(size_t)i * stride]);
    m = lr_block_max(m, red);
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += LR_THREADS) {
        const float e = expf(src[(size_t)i * stride] - m);
        dst[i] = e;
        s += e;
    }
    s = lr_block_sum(s, red);
    for (int i = threadIdx.x; i < n; i += LR_THREADS) dst[i] = dst[i] / s;
    __syncthreads();
}

// Rank one system: on entry sm.K[0..n) = its (softmaxed) scores in doc order, sm.pt = softmaxed labels.
// On exit K = scores in rank order, I = permutation, Y = labels in rank order, G = gains in rank order.
__device__ __forceinline__ void lr_rank_system(LrSmem& sm, int n, int len, float eps, float* red) {
    const float ninf = __int_as_float(0xff800000);
    block_sort_desc<false, LR_THREADS>(sm.K, sm.I, sm.K0, n, len, TIES_STDSORT);
    for (int i = threadIdx.x; i < len; i += LR_THREADS) sm.G[i] = (i < n) ? sm.pt[i] : ninf;
    block_bitonic_desc(sm.G, len);                                   // ideal label order (values only)
    float part = 0.f;
    for (int r = threadIdx.x; r < n; r += LR_THREADS) part += (exp2f(fmaxf(sm.G[r], 0.f)) - 1.0f) * sm.invD[r];
    const float maxdcg = fmaxf(lr_block_sum(part, red), eps);        // lambdaL.py:39 with k = None
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += LR_THREADS) {
        const float y = sm.pt[sm.I[r]];
        sm.Y[r] = y;
        sm.G[r] = (exp2f(fmaxf(y, 0.f)) - 1.0f) / maxdcg;            // :40
    }
    __syncthreads();
}

template <int SCHEME>
__device__ __forceinline__ void lr_colsums(const LrSmem& sm, int n, float mu, float sig_scale, float log_eps2, float* out) {
    for (int j = threadIdx.x; j < n; j += LR_THREADS) {
        const float sj = sm.K[j];
        float acc = 0.f;
        for (int i = 0; i < n; ++i) {
            float l, dl;
            pair_logprob<true>(sm.K[i] - sj, pair_weight<SCHEME>(i, j, sm.G, sm.invD, sm.dtab, sm.Y, mu), sig_scale, log_eps2,
                         1.0f, l, dl);
            acc += l;
        }
        out[j] = acc;
    }
    __syncthreads();
}

// effectiveness of `cur` against `base` + (for the model) the derivative w.r.t. cur
__device__ __forceinline__ float lr_effect(const float* base, const float* cur, int n, int lt, float* dvec_out, float* red) {
    if (lt == 1) {
        float a = 0.f;
        for (int j = threadIdx.x; j < n; j += LR_THREADS) {
            const float d = cur[j] - base[j];
            a = fmaf(d, d, a);
            if (dvec_out) dvec_out[j] = 2.0f * d;
        }
        return lr_block_sum(a, red);
    }
    if (lt == 3) {
        float a = 0.f, c = 0.f;
        for (int j = threadIdx.x; j < n; j += LR_THREADS) a += cur[j], c += base[j];
        const float u = lr_block_sum(a, red) - lr_block_sum(c, red);
        if (dvec_out)
            for (int j = threadIdx.x; j < n; j += LR_THREADS) dvec_out[j] = 2.0f * u;
        return u * u;
    }
    float xy = 0.f, xx = 0.f, yy = 0.f;
    for (int j = threadIdx.x; j < n; j += LR_THREADS) {
        xy = fmaf(base[j], cur[j], xy), xx = fmaf(base[j], base[j], xx), yy = fmaf(cur[j], cur[j], yy);
    }
    xy = lr_block_sum(xy, red), xx = lr_block_sum(xx, red), yy = lr_block_sum(yy, red);
    const float nx = fmaxf(sqrtf(xx), 1e-8f), ny = fmaxf(sqrtf(yy), 1e-8f);
    const float c = xy / (nx * ny);
    if (dvec_out)
        for (int j = threadIdx.x; j < n; j += LR_THREADS) dvec_out[j] = base[j] / (nx * ny) - c * cur[j] / (ny * ny);
    return c;
}

__device__ __forceinline__ void lr_tables(LrSmem& sm, int n) {
    for (int r = threadIdx.x; r < n; r += LR_THREADS) sm.invD[r] = 1.0f / log2f(2.0f + (float)r);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += LR_THREADS) sm.dtab[t] = (t == 0) ? 0.f : fabsf(sm.invD[t - 1] - sm.invD[t]);
    __syncthreads();
}

template <int SCHEME>
__global__ void __launch_bounds__(LR_THREADS) lambda_risk_rows_kernel(const LambdaRiskParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[32];
    LrSmem sm(smem_raw, p.len);
    const int n = p.n;
    const int C = 1 + p.S + p.add_ideal;
    const float sig_scale = p.sigma * kLog2e;
    const float log_eps2 = log2f(p.eps);
    lr_tables(sm, n);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        lr_softmax(p.y_true + (size_t)b * n, 1, n, sm.pt, red);
        // the labels ranked by themselves: the "base" column sums (riskLosses.py:72)
        for (int i = threadIdx.x; i < n; i += LR_THREADS) sm.K[i] = sm.pt[i];
        __syncthreads();
        lr_rank_system(sm, n, p.len, p.eps, red);
        lr_colsums<SCHEME>(sm, n, p.mu, sig_scale, log_eps2, sm.base);
        for (int sys = 0; sys <= p.S; ++sys) {
            if (sys == 0) lr_softmax(p.y_pred + (size_t)b * n, 1, n, sm.K, red);
            else lr_softmax(p.y_base + (size_t)b * n * p.S + (sys - 1), p.S, n, sm.K, red);
            lr_rank_system(sm, n, p.len, p.eps, red);
            lr_colsums<SCHEME>(sm, n, p.mu, sig_scale, log_eps2, sm.cur);
            const float v = lr_effect(sm.base, sm.cur, n, p.lt, (sys == 0 && p.dvec) ? p.dvec + (size_t)b * n : nullptr, red);
            if (threadIdx.x == 0) p.mat[(size_t)b * C + sys] = v;
            __syncthreads();
        }
        if (p.add_ideal) {
            // transformation 1: (base - base)^2 = 0 (:104); 2: ones for geoRisk (:107), cosine(base, base) for zRisk (:227)
            float v = 0.f;
            if (p.lt == 2) v = p.ideal_ones ? 1.0f : lr_effect(sm.base, sm.base, n, 2, nullptr, red);
            if (threadIdx.x == 0) p.mat[(size_t)b * C + 1 + p.S] = v;
            __syncthreads();
        }
    }
}

// grad[b, :] = gq[b] * sum_j dvec[b, j] * d colsum_j / d y_pred[b, :]
template <int SCHEME>
__global__ void __launch_bounds__(LR_THREADS) lambda_risk_bwd_kernel(const LambdaRiskParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[32];
    LrSmem sm(smem_raw, p.len);
    const int n = p.n;
    const float sig_scale = p.sigma * kLog2e;
    const float log_eps2 = log2f(p.eps);
    const float grad_scale = p.sigma / kLn2;                          // d log2(sig^w) / d delta = w sigma (1 - sig) / ln 2
    lr_tables(sm, n);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        lr_softmax(p.y_true + (size_t)b * n, 1, n, sm.pt, red);
        lr_softmax(p.y_pred + (size_t)b * n, 1, n, sm.K, red);
        for (int i = threadIdx.x; i < n; i += LR_THREADS) sm.base[i] = sm.K[i];     // softmax(y_pred) in doc order
        __syncthreads();
        lr_rank_system(sm, n, p.len, p.eps, red);
        const float up = p.gq[b];
        for (int j = threadIdx.x; j < n; j += LR_THREADS) sm.cur[j] = up * p.dvec[(size_t)b * n + j];   // g_j (rank order)
        __syncthreads();
        // d/d s_k = sum_j g_j dl_kj - g_k sum_i dl_ik   (k as a row, then k as the column)
        for (int k = threadIdx.x; k < n; k += LR_THREADS) {
            const float sk = sm.K[k], gk = sm.cur[k];
            float acc = 0.f;
            for (int j = 0; j < n; ++j) {
                float l, dl;
                pair_logprob<true>(sk - sm.K[j], pair_weight<SCHEME>(k, j, sm.G, sm.invD, sm.dtab, sm.Y, p.mu), sig_scale,
                             log_eps2, 1.0f, l, dl);
                acc = fmaf(sm.cur[j], dl, acc);
                pair_logprob<true>(sm.K[j] - sk, pair_weight<SCHEME>(j, k, sm.G, sm.invD, sm.dtab, sm.Y, p.mu), sig_scale,
                             log_eps2, 1.0f, l, dl);
                acc = fmaf(-gk, dl, acc);
            }
            sm.K0[sm.I[k]] = acc * grad_scale;                        // gradient w.r.t. softmax(y_pred), doc order
        }
        __syncthreads();
        // softmax backward
        float dot = 0.f;
        for (int i = threadIdx.x; i < n; i += LR_THREADS) dot = fmaf(sm.K0[i], sm.base[i], dot);
        dot = lr_block_sum(dot, red);
        for (int i = threadIdx.x; i < n; i += LR_THREADS) p.grad[(size_t)b * n + i] = sm.base[i] * (sm.K0[i] - dot);
        __syncthreads();
    }
}

// lambdaMask: the [n, n] matrix of pair log-probabilities of every slate (rank order) and the pair mask
struct PairMatrixParams {
    const float* y_pred;
    const float* y_true;
    int B, n, len, k;
    float eps, pad, sigma, mu;
    int natural_log;
    float* logp;           // [B, n, n]
    unsigned char* mask;   // [B, n, n] or null
    const float* glogp;    // backward: [B, n, n] upstream d L / d logp (rank order)
    float* grad;           // backward: [B, n] d L / d y_pred
};

// ranking prologue of lambdaMask (lambdaL.py:10-40) for slate b: scores / labels with the padded docs at -inf, sort with
// torch.sort's tie order, ideal DCG@k, labels and gains in rank order.  On exit K = scores in rank order, I = permutation,
// Y = labels in rank order, G = gains / maxDCG.  k < 0 (or >= n): no truncation; k = 0: maxDCG = eps (an empty sum).
__device__ __forceinline__ float pm_rank_prologue(LrSmem& sm, const PairMatrixParams& p, int b, float* red) {
    const int n = p.n;
    const float ninf = __int_as_float(0xff800000);
    const float* yp = p.y_true + (size_t)b * n;
    for (int i = threadIdx.x; i < n; i += LR_THREADS) {
        const bool pad = (yp[i] == p.pad);
        sm.K[i] = pad ? ninf : p.y_pred[(size_t)b * n + i];
        sm.pt[i] = pad ? ninf : yp[i];
    }
    __syncthreads();
    block_sort_desc<false, LR_THREADS>(sm.K, sm.I, sm.K0, n, p.len, TIES_STDSORT);
    for (int i = threadIdx.x; i < p.len; i += LR_THREADS) sm.G[i] = (i < n) ? sm.pt[i] : ninf;
    block_bitonic_desc(sm.G, p.len);
    const int kk = (p.k >= 0 && p.k < n) ? p.k : n;
    float part = 0.f;
    for (int r = threadIdx.x; r < kk; r += LR_THREADS) part += (exp2f(fmaxf(sm.G[r], 0.f)) - 1.0f) * sm.invD[r];
    const float maxdcg = fmaxf(lr_block_sum(part, red), p.eps);
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += LR_THREADS) {
        const float y = sm.pt[sm.I[r]];
        sm.Y[r] = y;
        sm.G[r] = (exp2f(fmaxf(y, 0.f)) - 1.0f) / maxdcg;
    }
    __syncthreads();
    return maxdcg;
}

template <int SCHEME>
__global__ void __launch_bounds__(LR_THREADS) lambda_pair_matrix_kernel(const PairMatrixParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[32];
    LrSmem sm(smem_raw, p.len);
    const int n = p.n;
    const float ninf = __int_as_float(0xff800000);
    const float sig_scale = p.sigma * kLog2e;
    const float log_eps2 = log2f(p.eps);
    const float out_scale = p.natural_log ? kLn2 : 1.0f;
    lr_tables(sm, n);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float maxdcg = pm_rank_prologue(sm, p, b, red);
        const int kk = (p.k >= 0 && p.k < n) ? p.k : n;
        (void)maxdcg;
        for (int idx = threadIdx.x; idx < n * n; idx += LR_THREADS) {
            const int i = idx / n, j = idx - i * n;
            float l, dl;
            pair_logprob<true>(sm.K[i] - sm.K[j], pair_weight<SCHEME>(i, j, sm.G, sm.invD, sm.dtab, sm.Y, p.mu), sig_scale,
                         log_eps2, out_scale, l, dl);
            p.logp[((size_t)b * n + i) * n + j] = l;
            if (p.mask) {
                const float yd = sm.Y[i] - sm.Y[j];
                bool ok = (yd == yd) && (fabsf(yd) != __int_as_float(0x7f800000));       // isfinite (lambdaL.py:24)
                if (SCHEME != SCHEME_NDCG1) ok = ok && (yd > 0.f);                        // :26-27
                ok = ok && (i < kk) && (j < kk);                                          // :29-30
                p.mask[((size_t)b * n + i) * n + j] = ok ? 1 : 0;
            }
        }
        __syncthreads();
    }
}

// backward of lambdaMask w.r.t. y_pred: logp[i, j] depends on the scores through delta = s_i - s_j only (the weights are
// functions of the labels and the rank positions, lambdaL.py:44-47 carry no gradient), so with u = glogp * d logp / d delta
//   d L / d s_(rank r) = sum_j u[r, j] - sum_i u[i, r];   padded docs get 0 (their scores are overwritten, :13).
// Thread r owns rank r.
template <int SCHEME>
__global__ void __launch_bounds__(LR_THREADS) lambda_pair_matrix_bwd_kernel(const PairMatrixParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[32];
    LrSmem sm(smem_raw, p.len);
    const int n = p.n;
    const float ninf = __int_as_float(0xff800000);
    const float sig_scale = p.sigma * kLog2e;
    const float log_eps2 = log2f(p.eps);
    const float out_scale = p.natural_log ? kLn2 : 1.0f;
    const float dscale = p.sigma * (p.natural_log ? 1.0f : 1.0f / kLn2);    // d log_b(sig^w) / d delta = w sigma (1 - sig) / ln b
    lr_tables(sm, n);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        pm_rank_prologue(sm, p, b, red);
        const float* g = p.glogp + (size_t)b * n * n;
        for (int r = threadIdx.x; r < n; r += LR_THREADS) {
            float acc = 0.f;
            const float sr = sm.K[r];
            if (sr != ninf) {
                for (int j = 0; j < n; ++j) {
                    float l, dl;
                    pair_logprob<true>(sr - sm.K[j], pair_weight<SCHEME>(r, j, sm.G, sm.invD, sm.dtab, sm.Y, p.mu), sig_scale,
                                       log_eps2, out_scale, l, dl);
                    acc = fmaf(g[(size_t)r * n + j], dl, acc);
                    pair_logprob<true>(sm.K[j] - sr, pair_weight<SCHEME>(j, r, sm.G, sm.invD, sm.dtab, sm.Y, p.mu), sig_scale,
                                       log_eps2, out_scale, l, dl);
                    acc = fmaf(-g[(size_t)j * n + r], dl, acc);
                }
            }
            p.grad[(size_t)b * n + sm.I[r]] = acc * dscale;
        }
        __syncthreads();
    }
}

template <typename Kern>
static int lr_prepare(Kern kern, size_t smem) {
    if (smem > 48 * 1024) LTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

template <int SCHEME>
static int lr_launch(int which, const LambdaRiskParams& p, size_t smem, int grid, cudaStream_t st) {
    if (which == 0) {
        if (int rc = lr_prepare(lambda_risk_rows_kernel<SCHEME>, smem)) return rc;
        lambda_risk_rows_kernel<SCHEME><<<grid, LR_THREADS, smem, st>>>(p);
    } else {
        if (int rc = lr_prepare(lambda_risk_bwd_kernel<SCHEME>, smem)) return rc;
        lambda_risk_bwd_kernel<SCHEME><<<grid, LR_THREADS, smem, st>>>(p);
    }
    return 0;
}

static int lr_dispatch(int scheme, int which, const LambdaRiskParams& p, size_t smem, int grid, cudaStream_t st) {
    switch (scheme) {
        case SCHEME_NONE: return lr_launch<SCHEME_NONE>(which, p, smem, grid, st);
        case SCHEME_NDCG1: return lr_launch<SCHEME_NDCG1>(which, p, smem, grid, st);
        case SCHEME_NDCG2: return lr_launch<SCHEME_NDCG2>(which, p, smem, grid, st);
        case SCHEME_LAMBDARANK: return lr_launch<SCHEME_LAMBDARANK>(which, p, smem, grid, st);
        case SCHEME_NDCG2PP: return lr_launch<SCHEME_NDCG2PP>(which, p, smem, grid, st);
        case SCHEME_GTDIFF: return lr_launch<SCHEME_GTDIFF>(which, p, smem, grid, st);
        default: return lr_launch<SCHEME_GTDIFF_POW>(which, p, smem, grid, st);
    }
}

template <int SCHEME>
static int pm_launch(const PairMatrixParams& p, size_t smem, int grid, cudaStream_t st) {
    if (p.grad) {
        if (int rc = lr_prepare(lambda_pair_matrix_bwd_kernel<SCHEME>, smem)) return rc;
        lambda_pair_matrix_bwd_kernel<SCHEME><<<grid, LR_THREADS, smem, st>>>(p);
        return 0;
    }
    if (int rc = lr_prepare(lambda_pair_matrix_kernel<SCHEME>, smem)) return rc;
    lambda_pair_matrix_kernel<SCHEME><<<grid, LR_THREADS, smem, st>>>(p);
    return 0;
}

static int pm_dispatch(int scheme, const PairMatrixParams& p, cudaStream_t st) {
    const size_t smem = (size_t)p.len * LR_ARRAYS * sizeof(float);
    const int cap = device_sm_count() * 8;
    const int grid = p.B < cap ? p.B : cap;
    int rc;
    switch (scheme) {
        case SCHEME_NONE: rc = pm_launch<SCHEME_NONE>(p, smem, grid, st); break;
        case SCHEME_NDCG1: rc = pm_launch<SCHEME_NDCG1>(p, smem, grid, st); break;
        case SCHEME_NDCG2: rc = pm_launch<SCHEME_NDCG2>(p, smem, grid, st); break;
        case SCHEME_LAMBDARANK: rc = pm_launch<SCHEME_LAMBDARANK>(p, smem, grid, st); break;
        case SCHEME_NDCG2PP: rc = pm_launch<SCHEME_NDCG2PP>(p, smem, grid, st); break;
        case SCHEME_GTDIFF: rc = pm_launch<SCHEME_GTDIFF>(p, smem, grid, st); break;
        default: rc = pm_launch<SCHEME_GTDIFF_POW>(p, smem, grid, st); break;
    }
    if (rc) return rc;
    return (int)cudaGetLastError();
}

static int lr_len(int n) {
    int len = 32;
    while (len < n) len <<= 1;
    return len;
}

}  // namespace ltr

LTR_API int ltr_risk_lambda_rows(const float* y_pred, const float* y_true, const float* y_baselines, int B, int n, int S,
                                 int transformation, int add_ideal, int ideal_ones, int scheme, float* mat, float* dvec,
                                 void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !mat || B <= 0 || n <= 0 || S < 0 || (S > 0 && !y_baselines) || scheme < 0 || scheme > 6)
        return set_error(LTR_EINVAL, "ltr_risk_lambda_rows: bad arguments");
    if (transformation < 1 || transformation > 3) return set_error(LTR_EINVAL, "ltr_risk_lambda_rows: transformation must be 1..3");
    if (n > LR_MAX_N) return set_error(LTR_ETOOBIG, "ltr_risk_lambda_rows: slate too long");
    LambdaRiskParams p{y_pred, y_true, y_baselines, B, n, lr_len(n), S, transformation, add_ideal ? 1 : 0, ideal_ones,
                       1e-10f, 1.0f, 10.0f, mat, dvec, nullptr, nullptr};       // lambdaMask's defaults (riskLosses.py:72)
    const size_t smem = (size_t)p.len * LR_ARRAYS * sizeof(float);
    const int cap = device_sm_count() * 8;
    int rc = lr_dispatch(scheme, 0, p, smem, B < cap ? B : cap, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_risk_lambda_rows") : 0;
}

LTR_API int ltr_risk_lambda_rows_bwd(const float* y_pred, const float* y_true, int B, int n, int scheme, const float* gq,
                                     const float* dvec, float* grad, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !gq || !dvec || !grad || B <= 0 || n <= 0 || scheme < 0 || scheme > 6)
        return set_error(LTR_EINVAL, "ltr_risk_lambda_rows_bwd: bad arguments");
    if (n > LR_MAX_N) return set_error(LTR_ETOOBIG, "ltr_risk_lambda_rows_bwd: slate too long");
    LambdaRiskParams p{y_pred, y_true, nullptr, B, n, lr_len(n), 0, 1, 0, 0, 1e-10f, 1.0f, 10.0f, nullptr,
                       const_cast<float*>(dvec), gq, grad};
    const size_t smem = (size_t)p.len * LR_ARRAYS * sizeof(float);
    const int cap = device_sm_count() * 8;
    int rc = lr_dispatch(scheme, 1, p, smem, B < cap ? B : cap, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_risk_lambda_rows_bwd") : 0;
}

LTR_API int ltr_lambda_pair_matrix(const float* y_pred, const float* y_true, int B, int n, float eps, float pad_value,
                                   int scheme, int k, float sigma, float mu, int natural_log, float* logp,
                                   unsigned char* mask, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !logp || B <= 0 || n <= 0 || scheme < 0 || scheme > 6)
        return set_error(LTR_EINVAL, "ltr_lambda_pair_matrix: bad arguments");
    if (n > LR_MAX_N) return set_error(LTR_ETOOBIG, "ltr_lambda_pair_matrix: slate too long");
    PairMatrixParams p{y_pred, y_true, B, n, lr_len(n), k, eps, pad_value, sigma, mu, natural_log, logp, mask, nullptr, nullptr};
    const int rc = pm_dispatch(scheme, p, reinterpret_cast<cudaStream_t>(stream));
    return rc ? set_cuda_error(rc, "ltr_lambda_pair_matrix") : 0;
}

LTR_API int ltr_lambda_pair_matrix_bwd(const float* y_pred, const float* y_true, int B, int n, float eps, float pad_value,
                                       int scheme, int k, float sigma, float mu, int natural_log, const float* grad_logp,
                                       float* grad, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !grad_logp || !grad || B <= 0 || n <= 0 || scheme < 0 || scheme > 6)
        return set_error(LTR_EINVAL, "ltr_lambda_pair_matrix_bwd: bad arguments");
    if (n > LR_MAX_N) return set_error(LTR_ETOOBIG, "ltr_lambda_pair_matrix_bwd: slate too long");
    PairMatrixParams p{y_pred, y_true, B, n, lr_len(n), k, eps, pad_value, sigma, mu, natural_log, nullptr, nullptr, grad_logp, grad};
    const int rc = pm_dispatch(scheme, p, reinterpret_cast<cudaStream_t>(stream));
    return rc ? set_cuda_error(rc, "ltr_lambda_pair_matrix_bwd") : 0;
}
