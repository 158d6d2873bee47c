// K4 / K5 / K8: kernels that need the exact rank order of a slate.
//   ltr_sort_desc        torch.sort(descending=True) / argsort permutation with the reference's tie order
//   ltr_ndcg_at_k        utils/metrics.py:83-104 torchNdcg  (fp32, exponential gains, all k of a sweep in one read)
//   ltr_ndcg_at_k_f64    utils/metrics.py:48-80  dcg / ndcg / mNdcg (fp64, linear or exponential gains)
//   ltr_exact_ndcg       losses/exactNDCG.py:5-29 ndcgLoss (forward only: the reference gives y_pred no gradient)
//   ltr_order_score      losses/orderScore.py:5-24 orderScoreLoss, fused forward + backward
// One CTA sorts one slate in shared memory (sort_common.cuh).  HBM traffic is 8 n bytes per query
// (scores + labels) plus the outputs; the sort and the top-k sums never leave the SM.
#include "ltr_host.h"
#include "sort_common.cuh"

namespace ltr {

constexpr int RANK_THREADS = 128;
constexpr int RANK_MAX_N = 4096;
constexpr int RANK_MAX_K = 32;

__device__ __forceinline__ float rk_block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
    return t;
}
__device__ __forceinline__ float rk_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// shared-memory carve-up for one slate: K[len] I[len] K0[len] L[len]
struct RankSmem {
    float* K;
    int* I;
    float* K0;
    float* L;
    __device__ RankSmem(void* base, int len) {
        K = reinterpret_cast<float*>(base);
        I = reinterpret_cast<int*>(K + len);
        K0 = reinterpret_cast<float*>(I + len);
        L = K0 + len;
    }
};

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RANK_THREADS) sort_desc_kernel(const float* keys, int B, int n, int len, int tie_rule,
                                                                 int* perm, float* sorted) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RankSmem sm(smem_raw, len);
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sm.K[i] = keys[(size_t)b * n + i];
        __syncthreads();
        block_sort_desc(sm.K, sm.I, sm.K0, n, len, tie_rule);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (perm) perm[(size_t)b * n + i] = sm.I[i];
            if (sorted) sorted[(size_t)b * n + i] = sm.K[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// NDCG@k for every k of `ks` (k <= 0: no truncation).  F64 = false: torchNdcg arithmetic (fp32, gains
// 2^l - 1, discount table disc32[r] = float(log2(2 + r)), sequential accumulation, 0/0 = NaN).
// F64 = true: metrics.dcg arithmetic (fp64, linear or exponential gains, disc64[r] = log2(2 + r),
// k clipped to n, ideal 0 -> `empty_value`).
struct NdcgParams {
    const float* y_true;
    const float* y_pred;
    int B, n, len, nk, tie_rule, exp_gains;
    int ks[RANK_MAX_K];
    const float* disc32;
    const double* disc64;
    double empty_value;
    float* out32;
    double* out64;
};

template <bool F64>
__global__ void __launch_bounds__(RANK_THREADS) ndcg_kernel(const NdcgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RankSmem sm(smem_raw, p.len);
    const float ninf = __int_as_float(0xff800000);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* yt = p.y_true + (size_t)b * p.n;
        for (int i = threadIdx.x; i < p.n; i += blockDim.x) sm.K[i] = p.y_pred[(size_t)b * p.n + i];
        __syncthreads();
        block_sort_desc(sm.K, sm.I, sm.K0, p.n, p.len, p.tie_rule);
        // labels in predicted order (K0) and in ideal order (L; values only, so ties are immaterial)
        for (int i = threadIdx.x; i < p.len; i += blockDim.x) {
            sm.L[i] = (i < p.n) ? yt[i] : ninf;
            if (i < p.n) sm.K0[i] = yt[sm.I[i]];
        }
        block_bitonic_desc(sm.L, p.len);
        if ((int)threadIdx.x < p.nk) {
            int k = p.ks[threadIdx.x];
            if (k <= 0 || k > p.n) k = p.n;
            if (!F64) {
                float got = 0.f, ideal = 0.f;
                for (int r = 0; r < k; ++r) {
                    const float d = p.disc32[r];
                    got += (exp2f(sm.K0[r]) - 1.0f) / d;
                    ideal += (exp2f(sm.L[r]) - 1.0f) / d;
                }
                p.out32[(size_t)b * p.nk + threadIdx.x] = got / ideal;
            } else {
                double got = 0.0, ideal = 0.0;
                for (int r = 0; r < k; ++r) {
                    const double d = p.disc64[r];
                    const double lg = (double)sm.K0[r], li = (double)sm.L[r];
                    got += (p.exp_gains ? exp2(lg) - 1.0 : lg) / d;
                    ideal += (p.exp_gains ? exp2(li) - 1.0 : li) / d;
                }
                p.out64[(size_t)b * p.nk + threadIdx.x] = (ideal == 0.0) ? p.empty_value : got / ideal;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// exactNDCG (MODE 0) and orderScore (MODE 1).  Both work on softmax(y_true) / softmax(y_pred).
struct RankLossParams {
    const float* y_true;
    const float* y_pred;
    int B, n, len;
    float* slate;     // [B] per-slate value
    float* grad;      // [B, n] (orderScore) or null
};

template <int MODE>
__global__ void __launch_bounds__(RANK_THREADS) rank_loss_kernel(const RankLossParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[32];
    RankSmem sm(smem_raw, p.len);
    const float ninf = __int_as_float(0xff800000);
    const int n = p.n;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* yt = p.y_true + (size_t)b * n;
        const float* yp = p.y_pred + (size_t)b * n;
        // softmax statistics of both rows
        float mt = ninf, mp = ninf;
        for (int i = threadIdx.x; i < n; i += blockDim.x) mt = fmaxf(mt, yt[i]), mp = fmaxf(mp, yp[i]);
        mt = rk_block_max(mt, red);
        mp = rk_block_max(mp, red);
        float st = 0.f, sp = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) st += expf(yt[i] - mt), sp += expf(yp[i] - mp);
        st = rk_block_sum(st, red);
        sp = rk_block_sum(sp, red);
        if (MODE == 0) {
            // order by softmax(y_pred) (exactNDCG.py:12), gains softmax(y_true), discount ln(r) + 1 (:17)
            for (int i = threadIdx.x; i < n; i += blockDim.x) sm.K[i] = expf(yp[i] - mp) / sp;
            __syncthreads();
            block_sort_desc(sm.K, sm.I, sm.K0, n, p.len, TIES_STDSORT);
            for (int i = threadIdx.x; i < p.len; i += blockDim.x) sm.L[i] = (i < n) ? expf(yt[i] - mt) / st : ninf;
            __syncthreads();
            float dcg = 0.f;
            for (int r = threadIdx.x; r < n; r += blockDim.x) dcg += (expf(yt[sm.I[r]] - mt) / st) / (logf((float)(r + 1)) + 1.0f);
            dcg = rk_block_sum(dcg, red);
            block_bitonic_desc(sm.L, p.len);
            float idcg = 0.f;
            for (int r = threadIdx.x; r < n; r += blockDim.x) idcg += sm.L[r] / (logf((float)(r + 1)) + 1.0f);
            idcg = rk_block_sum(idcg, red);
            if (threadIdx.x == 0) p.slate[b] = -(dcg / idcg);
        } else {
            // order by softmax(y_true) (orderScore.py:14), value sum_r (-n + 2r) log softmax(y_pred)[order_r]
            for (int i = threadIdx.x; i < n; i += blockDim.x) sm.K[i] = expf(yt[i] - mt) / st;
            __syncthreads();
            block_sort_desc(sm.K, sm.I, sm.K0, n, p.len, TIES_STDSORT);
            const float lse = mp + logf(sp);
            const float invB = 1.0f / (float)p.B;
            float acc = 0.f;
            for (int r = threadIdx.x; r < n; r += blockDim.x) {
                const int doc = sm.I[r];
                const float ramp = (float)(2 * r - n);
                const float s = yp[doc];
                acc += ramp * (s - lse);
                if (p.grad) p.grad[(size_t)b * n + doc] = -invB * (ramp + (float)n * (expf(s - mp) / sp));
            }
            acc = rk_block_sum(acc, red);
            if (threadIdx.x == 0) p.slate[b] = -acc;
        }
        __syncthreads();
    }
}

// out[0] = mean of v[0..B) with a fixed reduction shape (one CTA)
__global__ void __launch_bounds__(256) mean_kernel(const float* v, int B, float* out) {
    __shared__ double red[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) acc += (double)v[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[0] = (float)(t / (double)B);
    }
}

static int rank_len(int n) {
    int len = 1;
    while (len < n) len <<= 1;
    return len < 32 ? 32 : len;
}
static size_t rank_smem(int len) { return (size_t)len * 16; }

template <typename Kern>
static int rank_prepare(Kern kern, size_t smem) {
    if (smem > 48 * 1024) LTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}
static int rank_grid(int B) {
    const int cap = device_sm_count() * 8;
    return B < cap ? B : cap;
}

}  // namespace ltr

LTR_API int ltr_sort_desc(const float* keys, int B, int n, int tie_rule, int* perm, float* sorted_keys, void* stream) {
    using namespace ltr;
    if (!keys || B <= 0 || n <= 0 || tie_rule < 0 || tie_rule > 2)
        return set_error(LTR_EINVAL, "ltr_sort_desc: bad arguments");
    if (n > RANK_MAX_N) return set_error(LTR_ETOOBIG, "ltr_sort_desc: slate too long");
    const int len = rank_len(n);
    const size_t smem = rank_smem(len);
    if (int rc = rank_prepare(sort_desc_kernel, smem)) return rc;
    sort_desc_kernel<<<rank_grid(B), RANK_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(keys, B, n, len, tie_rule,
                                                                                                  perm, sorted_keys);
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_sort_desc") : 0;
}

static int ndcg_common(ltr::NdcgParams& p, const int* ks, int nk, bool f64, void* stream) {
    using namespace ltr;
    if (!p.y_true || !p.y_pred || p.B <= 0 || p.n <= 0 || nk <= 0 || nk > RANK_MAX_K || !ks)
        return set_error(LTR_EINVAL, "ltr_ndcg_at_k: bad arguments");
    if (p.n > RANK_MAX_N) return set_error(LTR_ETOOBIG, "ltr_ndcg_at_k: slate too long");
    p.len = rank_len(p.n);
    p.nk = nk;
    for (int j = 0; j < nk; ++j) p.ks[j] = ks[j];
    const size_t smem = rank_smem(p.len);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (f64) {
        if (int rc = rank_prepare(ndcg_kernel<true>, smem)) return rc;
        ndcg_kernel<true><<<rank_grid(p.B), RANK_THREADS, smem, st>>>(p);
    } else {
        if (int rc = rank_prepare(ndcg_kernel<false>, smem)) return rc;
        ndcg_kernel<false><<<rank_grid(p.B), RANK_THREADS, smem, st>>>(p);
    }
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_ndcg_at_k") : 0;
}

LTR_API int ltr_ndcg_at_k(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                          const float* disc32, float* out, void* stream) {
    ltr::NdcgParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.B = Q; p.n = n; p.tie_rule = tie_rule; p.disc32 = disc32; p.out32 = out;
    if (!disc32 || !out) return ltr::set_error(LTR_EINVAL, "ltr_ndcg_at_k: null table / output");
    return ndcg_common(p, ks, nk, false, stream);
}

LTR_API int ltr_ndcg_at_k_f64(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                              int exponential_gains, double empty_value, const double* disc64, double* out,
                              void* stream) {
    ltr::NdcgParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.B = Q; p.n = n; p.tie_rule = tie_rule; p.disc64 = disc64; p.out64 = out;
    p.exp_gains = exponential_gains; p.empty_value = empty_value;
    if (!disc64 || !out) return ltr::set_error(LTR_EINVAL, "ltr_ndcg_at_k_f64: null table / output");
    return ndcg_common(p, ks, nk, true, stream);
}

static int rank_loss_common(int mode, const float* y_true, const float* y_pred, int B, int n, float* slate, float* loss,
                            float* grad, void* stream, const char* what) {
    using namespace ltr;
    if (!y_true || !y_pred || !slate || !loss || B <= 0 || n <= 0) return set_error(LTR_EINVAL, what);
    if (n > RANK_MAX_N) return set_error(LTR_ETOOBIG, what);
    RankLossParams p{y_true, y_pred, B, n, rank_len(n), slate, grad};
    const size_t smem = rank_smem(p.len);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (mode == 0) {
        if (int rc = rank_prepare(rank_loss_kernel<0>, smem)) return rc;
        rank_loss_kernel<0><<<rank_grid(B), RANK_THREADS, smem, st>>>(p);
    } else {
        if (int rc = rank_prepare(rank_loss_kernel<1>, smem)) return rc;
        rank_loss_kernel<1><<<rank_grid(B), RANK_THREADS, smem, st>>>(p);
    }
    mean_kernel<<<1, 256, 0, st>>>(slate, B, loss);
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, what) : 0;
}

LTR_API int ltr_exact_ndcg(const float* y_true, const float* y_pred, int B, int n, float* slate_value, float* loss,
                           void* stream) {
    return rank_loss_common(0, y_true, y_pred, B, n, slate_value, loss, nullptr, stream, "ltr_exact_ndcg");
}

LTR_API int ltr_order_score_fwdbwd(const float* y_true, const float* y_pred, int B, int n, float* slate_value, float* loss,
                                   float* grad, void* stream) {
    return rank_loss_common(1, y_true, y_pred, B, n, slate_value, loss, grad, stream, "ltr_order_score_fwdbwd");
}
