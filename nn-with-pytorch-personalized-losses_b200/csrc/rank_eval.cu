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
    int kmax;                // largest k of the sweep (top-k fast path)
    int* flags;              // [1 + B] scratch: flags[0] = number of queries the top-k path left to the full sort
                             // (a tie among their first kmax + 1 ranks), flags[1..] = their indices
};

template <bool F64>
__global__ void __launch_bounds__(RANK_THREADS) ndcg_kernel(const NdcgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RankSmem sm(smem_raw, p.len);
    const float ninf = __int_as_float(0xff800000);
    // with the top-k fast path in front, only the queries it appended to the flagged list are sorted here
    const int work = p.flags ? p.flags[0] : p.B;
    for (int w = blockIdx.x; w < work; w += gridDim.x) {
        const int b = p.flags ? p.flags[1 + w] : w;
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
// Top-k fast path of the NDCG sweep (all k <= 31): one WARP per query, the row in registers, no sort.
//   predicted order: kmax rounds of "lane-local max (a max tree, no index), REDUX.MAX over the warp, retire
//     the element(s) equal to it".  If a round retires more than one element, or two consecutive rounds give
//     equal keys, a tie group reaches into the first kmax ranks: with LTR_TIES_STDSORT (and n > 16) the query is appended
//     to the flagged list and left to the full-sort kernel; with the index-ordered tie rules the round falls
//     back to an exact arg-max on (key, index).  A tie group entirely below rank kmax cannot change the top-k ranks.
//   ideal order: labels that are integers 0..7 (MSLR / Istella: 0..4) are counted per level with ballots and rank r
//     reads its label off the cumulative counts; other labels take top-k rounds of their own (values only).
//   DCG sums: lane r holds the term of rank r; the running sums are formed in rank order (the reference
//     accumulates sequentially in fp32 / fp64) and written where r + 1 equals a k of the sweep.
// 8 n bytes of HBM traffic per query.
__device__ __forceinline__ int topk_key(float f) {
    if (f != f) return 0x7fffffff;                                // NaN first (at::native KeyValueCompDesc)
    return float_order_key(f + 0.0f);                             // -0 -> +0 so that key equality == float equality
}

template <int ITEMS>
__device__ __forceinline__ int tree_max(const int (&v)[ITEMS]) {
    int m[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) m[k] = v[k];
#pragma unroll
    for (int w = ITEMS / 2; w > 0; w >>= 1) {
#pragma unroll
        for (int k = 0; k < w; ++k) m[k] = max(m[k], m[k + w]);
    }
    return m[0];
}

template <int ITEMS, bool F64>
__global__ void __launch_bounds__(256) ndcg_topk_kernel(const NdcgParams p) {
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * 8;
    const int TAKEN = (int)0x80000000;
    const bool idx_desc = (p.tie_rule == TIES_REVERSED);
    const bool want_std = (p.tie_rule == TIES_STDSORT) && p.n > 16;
    int myk = -1;                                                 // lane j writes the j-th k of the sweep
    for (int j = 0; j < p.nk; ++j) myk = (lane == j) ? p.ks[j] : myk;
    for (int b = gwarp; b < p.B; b += nwarps) {
        const float* sp = p.y_pred + (size_t)b * p.n;
        const float* yp = p.y_true + (size_t)b * p.n;
        int ks_[ITEMS];
        float yv[ITEMS];
        bool small_int = true;
        int lmax = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = lane + 32 * k;
            const bool in = i < p.n;
            const float s = in ? ld_stream(sp + i) : 0.f;
            yv[k] = in ? ld_stream(yp + i) : 0.f;
            ks_[k] = in ? topk_key(s) : TAKEN;
            const int lv = (int)fminf(fmaxf(yv[k], -1.f), 8.f);
            small_int = small_int && (!in || ((float)lv == yv[k] && lv >= 0 && lv <= 7));
            lmax = max(lmax, lv);
        }
        small_int = __all_sync(LTR_FULL_MASK, small_int);
        lmax = __reduce_max_sync(LTR_FULL_MASK, lmax);

        // ---- predicted order: label at rank r ends up in lane r --------------------------------------
        float lab_pred = 0.f;
        int prev_key = 0x7ffffffe, tie = 0;
        for (int r = 0; r < p.kmax; ++r) {
            const int wk = __reduce_max_sync(LTR_FULL_MASK, tree_max(ks_));
            int hits = 0;
            float mine = 0.f;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const bool hit = (ks_[k] == wk) && (wk != TAKEN);
                hits += hit ? 1 : 0;
                mine = hit ? yv[k] : mine;
            }
            const unsigned owners = __ballot_sync(LTR_FULL_MASK, hits > 0);
            const bool multi = (__popc(owners) > 1) || __any_sync(LTR_FULL_MASK, hits > 1);
            int widx = -1;                                        // exact winner index, only needed on a tie
            if (multi || wk == prev_key) {
                if (want_std) { tie = 1; break; }
                int cand = idx_desc ? -1 : 0x7fffffff;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k)
                    if (ks_[k] == wk) cand = idx_desc ? max(cand, lane + 32 * k) : min(cand, lane + 32 * k);
                widx = idx_desc ? __reduce_max_sync(LTR_FULL_MASK, cand) : __reduce_min_sync(LTR_FULL_MASK, cand);
            }
            int owner = __ffs(owners) - 1;
            if (widx >= 0) {
                owner = widx & 31;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) {
                    if (lane + 32 * k == widx) { mine = yv[k]; ks_[k] = TAKEN; }
                }
            } else {
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) ks_[k] = (ks_[k] == wk) ? TAKEN : ks_[k];
            }
            const float labw = __shfl_sync(LTR_FULL_MASK, mine, owner < 0 ? 0 : owner);
            if (lane == r) lab_pred = (wk == TAKEN) ? 0.f : labw;
            prev_key = (wk == TAKEN) ? 0x7ffffffe - r : wk;
        }
        if (tie) {                                                // left to the full sort (std::sort tie order)
            if (lane == 0) p.flags[1 + atomicAdd(p.flags, 1)] = b;
            continue;
        }

        // ---- ideal order -----------------------------------------------------------------------------
        float lab_ideal = 0.f;
        if (small_int) {
            int ge = 0;                                           // lane r: number of levels L >= 1 with #(label >= L) > r
            int cum = 0;
            for (int L = lmax; L >= 1; --L) {
                int c = 0;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k)
                    c += __popc(__ballot_sync(LTR_FULL_MASK, (lane + 32 * k < p.n) && yv[k] == (float)L));
                cum += c;                                         // docs with label >= L
                ge += (cum > lane) ? 1 : 0;
            }
            lab_ideal = (float)ge;                                // labels fill the ranks from the top level down
        } else {
            int kl_[ITEMS];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) kl_[k] = (lane + 32 * k < p.n) ? topk_key(yv[k]) : TAKEN;
            for (int r = 0; r < p.kmax; ++r) {
                const int wl = __reduce_max_sync(LTR_FULL_MASK, tree_max(kl_));
                bool have = false;
                float mine = 0.f;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k)
                    if (!have && kl_[k] == wl && wl != TAKEN) { have = true; mine = yv[k]; }
                const int owner = __ffs(__ballot_sync(LTR_FULL_MASK, have)) - 1;
                bool done = false;
#pragma unroll
                for (int k = 0; k < ITEMS; ++k)
                    if (lane == owner && !done && kl_[k] == wl) { kl_[k] = TAKEN; done = true; }
                const float labi = __shfl_sync(LTR_FULL_MASK, mine, owner < 0 ? 0 : owner);
                if (lane == r) lab_ideal = (owner < 0) ? 0.f : labi;
            }
        }

        // ---- DCG sums in rank order ------------------------------------------------------------------
        if (!F64) {
            const float d = (lane < p.kmax) ? p.disc32[lane] : 1.0f;
            const float tp = (exp2f(lab_pred) - 1.0f) / d, ti = (exp2f(lab_ideal) - 1.0f) / d;
            float got = 0.f, ideal = 0.f;
            for (int r = 0; r < p.kmax; ++r) {
                got += __shfl_sync(LTR_FULL_MASK, tp, r);
                ideal += __shfl_sync(LTR_FULL_MASK, ti, r);
                if (myk == r + 1) p.out32[(size_t)b * p.nk + lane] = got / ideal;
            }
        } else {
            const double d = (lane < p.kmax) ? p.disc64[lane] : 1.0;
            const double lp = (double)lab_pred, li = (double)lab_ideal;
            const double tp = (p.exp_gains ? exp2(lp) - 1.0 : lp) / d, ti = (p.exp_gains ? exp2(li) - 1.0 : li) / d;
            double got = 0.0, ideal = 0.0;
            for (int r = 0; r < p.kmax; ++r) {
                got += __shfl_sync(LTR_FULL_MASK, tp, r);
                ideal += __shfl_sync(LTR_FULL_MASK, ti, r);
                if (myk == r + 1) p.out64[(size_t)b * p.nk + lane] = (ideal == 0.0) ? p.empty_value : got / ideal;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Long rows (256 < n <= 1024): one CTA of 4 warps per query.  Every warp extracts the top-kmax (key, label)
// candidates of its 256-doc quarter with the rounds above (8 docs per lane, so registers stay low and many
// queries are resident per SM), the per-warp label histograms and candidates meet in shared memory, and warp 0
// runs the same rounds over the <= 4 kmax candidates.  Any tie met on the way, and labels that are not integers
// 0..7, send the query to the full-sort kernel.
template <bool F64>
__global__ void __launch_bounds__(128) ndcg_topk_wide_kernel(const NdcgParams p) {
    constexpr int ITEMS = 8;
    __shared__ int cand_key[4][32];
    __shared__ float cand_lab[4][32];
    __shared__ int hist[4][8];
    __shared__ int sh_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TAKEN = (int)0x80000000;
    int myk = -1;
    for (int j = 0; j < p.nk; ++j) myk = (lane == j) ? p.ks[j] : myk;
    if (tid == 0) sh_flag = 0;
    __syncthreads();
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* sp = p.y_pred + (size_t)b * p.n;
        const float* yp = p.y_true + (size_t)b * p.n;
        int ks_[ITEMS];
        float yv[ITEMS];
        bool small_int = true;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = warp * 256 + lane + 32 * k;
            const bool in = i < p.n;
            const float s = in ? ld_stream(sp + i) : 0.f;
            yv[k] = in ? ld_stream(yp + i) : 0.f;
            ks_[k] = in ? topk_key(s) : TAKEN;
            const int lv = (int)fminf(fmaxf(yv[k], -1.f), 8.f);
            small_int = small_int && (!in || ((float)lv == yv[k] && lv >= 0 && lv <= 7));
        }
        int bad = __all_sync(LTR_FULL_MASK, small_int) ? 0 : 1;
        // per-warp label histogram (levels 1..7)
        for (int L = 1; L <= 7; ++L) {
            int c = 0;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k)
                c += __popc(__ballot_sync(LTR_FULL_MASK, (warp * 256 + lane + 32 * k < p.n) && yv[k] == (float)L));
            if (lane == 0) hist[warp][L] = c;
        }
        // per-warp candidates
        int ckey = TAKEN;
        float clab = 0.f;
        for (int r = 0; r < p.kmax; ++r) {
            const int wk = __reduce_max_sync(LTR_FULL_MASK, tree_max(ks_));
            int hits = 0;
            float mine = 0.f;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const bool hit = (ks_[k] == wk) && (wk != TAKEN);
                hits += hit ? 1 : 0;
                mine = hit ? yv[k] : mine;
                ks_[k] = hit ? TAKEN : ks_[k];
            }
            const unsigned owners = __ballot_sync(LTR_FULL_MASK, hits > 0);
            if (__popc(owners) > 1 || __any_sync(LTR_FULL_MASK, hits > 1)) bad = 1;
            const float labw = __shfl_sync(LTR_FULL_MASK, mine, owners ? __ffs(owners) - 1 : 0);
            if (lane == r) { ckey = wk; clab = labw; }
        }
        cand_key[warp][lane] = (lane < p.kmax) ? ckey : TAKEN;
        cand_lab[warp][lane] = clab;
        if (bad && lane == 0) sh_flag = 1;       // (benign race: every writer stores 1)
        __syncthreads();
        if (warp == 0) {
            int tie = sh_flag;
            __syncwarp();
            if (lane == 0) sh_flag = 0;          // re-armed for the next query (ordered by the barrier below)
            int mk[4];
            float ml[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) mk[k] = cand_key[k][lane], ml[k] = cand_lab[k][lane];
            float lab_pred = 0.f;
            for (int r = 0; r < p.kmax && !tie; ++r) {
                const int wk = __reduce_max_sync(LTR_FULL_MASK, tree_max(mk));
                int hits = 0;
                float mine = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool hit = (mk[k] == wk) && (wk != TAKEN);
                    hits += hit ? 1 : 0;
                    mine = hit ? ml[k] : mine;
                    mk[k] = hit ? TAKEN : mk[k];
                }
                const unsigned owners = __ballot_sync(LTR_FULL_MASK, hits > 0);
                if (__popc(owners) > 1 || __any_sync(LTR_FULL_MASK, hits > 1)) tie = 1;
                const float labw = __shfl_sync(LTR_FULL_MASK, mine, owners ? __ffs(owners) - 1 : 0);
                if (lane == r) lab_pred = (wk == TAKEN) ? 0.f : labw;
            }
            if (tie) {
                if (lane == 0) p.flags[1 + atomicAdd(p.flags, 1)] = b;
            } else {
                int ge = 0, cum = 0;
                for (int L = 7; L >= 1; --L) {
                    cum += (hist[0][L] + hist[1][L]) + (hist[2][L] + hist[3][L]);
                    ge += (cum > lane) ? 1 : 0;
                }
                const float lab_ideal = (float)ge;
                if (!F64) {
                    const float d = (lane < p.kmax) ? p.disc32[lane] : 1.0f;
                    const float tp = (exp2f(lab_pred) - 1.0f) / d, ti = (exp2f(lab_ideal) - 1.0f) / d;
                    float got = 0.f, ideal = 0.f;
                    for (int r = 0; r < p.kmax; ++r) {
                        got += __shfl_sync(LTR_FULL_MASK, tp, r);
                        ideal += __shfl_sync(LTR_FULL_MASK, ti, r);
                        if (myk == r + 1) p.out32[(size_t)b * p.nk + lane] = got / ideal;
                    }
                } else {
                    const double d = (lane < p.kmax) ? p.disc64[lane] : 1.0;
                    const double lp = (double)lab_pred, li = (double)lab_ideal;
                    const double tp = (p.exp_gains ? exp2(lp) - 1.0 : lp) / d, ti = (p.exp_gains ? exp2(li) - 1.0 : li) / d;
                    double got = 0.0, ideal = 0.0;
                    for (int r = 0; r < p.kmax; ++r) {
                        got += __shfl_sync(LTR_FULL_MASK, tp, r);
                        ideal += __shfl_sync(LTR_FULL_MASK, ti, r);
                        if (myk == r + 1) p.out64[(size_t)b * p.nk + lane] = (ideal == 0.0) ? p.empty_value : got / ideal;
                    }
                }
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
    // effective k of every entry (k <= 0 or k > n: the whole slate, metrics.py:54-55,87-88)
    int kmax = 0;
    for (int j = 0; j < nk; ++j) {
        p.ks[j] = (ks[j] <= 0 || ks[j] > p.n) ? p.n : ks[j];
        kmax = p.ks[j] > kmax ? p.ks[j] : kmax;
    }
    p.kmax = kmax;
    const size_t smem = rank_smem(p.len);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (kmax <= 31 && p.n <= 1024 && p.flags) {
        // top-k fast path (one warp per query), then the full sort for the queries it flagged
        LTR_CUDA(cudaMemsetAsync(p.flags, 0, sizeof(int), st));
        int g = (p.B + 7) / 8;
        const int cap = device_sm_count() * 8;
        g = g < cap ? g : cap;
#define LTR_TOPK(ITEMS)                                                      \
        if (f64) ndcg_topk_kernel<ITEMS, true><<<g, 256, 0, st>>>(p);            \
        else ndcg_topk_kernel<ITEMS, false><<<g, 256, 0, st>>>(p)
        bool wide = false;
        if (p.n <= 128) { LTR_TOPK(4); }
        else if (p.n <= 256) { LTR_TOPK(8); }
        else {
            wide = true;                           // ties of any rule go to the full sort there
            const int gw = p.B < cap ? p.B : cap;
            if (f64) ndcg_topk_wide_kernel<true><<<gw, 128, 0, st>>>(p);
            else ndcg_topk_wide_kernel<false><<<gw, 128, 0, st>>>(p);
        }
#undef LTR_TOPK
        if (!wide && (p.tie_rule != TIES_STDSORT || p.n <= 16)) {
            const int rc0 = (int)cudaGetLastError();
            return rc0 ? set_cuda_error(rc0, "ltr_ndcg_at_k") : 0;
        }
    } else {
        p.flags = nullptr;
    }
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
                          const float* disc32, float* out, int* flags, void* stream) {
    ltr::NdcgParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.B = Q; p.n = n; p.tie_rule = tie_rule; p.disc32 = disc32; p.out32 = out;
    p.flags = flags;
    if (!disc32 || !out) return ltr::set_error(LTR_EINVAL, "ltr_ndcg_at_k: null table / output");
    return ndcg_common(p, ks, nk, false, stream);
}

LTR_API int ltr_ndcg_at_k_f64(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                              int exponential_gains, double empty_value, const double* disc64, double* out,
                              int* flags, void* stream) {
    ltr::NdcgParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.B = Q; p.n = n; p.tie_rule = tie_rule; p.disc64 = disc64; p.out64 = out;
    p.flags = flags;
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
