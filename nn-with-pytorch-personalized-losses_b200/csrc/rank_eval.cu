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
        block_sort_desc<false, RANK_THREADS>(sm.K, sm.I, sm.K0, n, len, tie_rule);
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
    int raw_dcg;             // F64 only: write the DCG itself instead of DCG / ideal DCG (metrics.dcg, :48-66)
    int ks[RANK_MAX_K];
    const float* disc32;
    const double* disc64;
    double empty_value;
    float* out32;
    double* out64;
    int kmax;                // largest k of the sweep (top-k fast path)
    int* flags;              // [1 + B] scratch: flags[1 + q] != 0 marks a query the top-k path left to the full sort
                             // (a tie among its first kmax + 1 ranks, a NaN score, an unusual label)
    float* partial;          // long rows in 128-doc chunks (ndcg_keys.cuh): [13][rows] words -- a chunk's 11 largest keys, its label
                             // histogram and its "not on the fast path" flag -- written by ndcg_key_kernel INSTEAD of its outputs
                             // (the kernel then sees every chunk as a row of 128 docs), read by ndcg_chunk_merge_kernel; else null
};

// one query: full sort with the caller's tie rule, labels in predicted and in ideal order, all k of the sweep
// known_tie: the caller has seen equal keys among the first kmax + 1 ranks.  have_hist: every label of the row is an
// integer 0..4 and level L occurs (hist >> 8 (L - 1)) & 255 times -- the ideal order is then written down directly.
template <bool F64, bool WARP = false>
__device__ __forceinline__ void ndcg_sorted_query(const NdcgParams& p, RankSmem& sm, int b, bool known_tie = false,
                                                  bool have_hist = false, unsigned hist = 0, bool count_hist = false) {
    const float ninf = __int_as_float(0xff800000);
    const float* yt = p.y_true + (size_t)b * p.n;
    for (int i = grp_tid<WARP>(); i < p.n; i += grp_size<WARP>()) sm.K[i] = p.y_pred[(size_t)b * p.n + i];
    grp_sync<WARP>();
    block_sort_desc<WARP, WARP ? 0 : RANK_THREADS>(sm.K, sm.I, sm.K0, p.n, p.len, p.tie_rule, p.kmax + 1,
                          WARP ? reinterpret_cast<int*>(sm.L + p.len) : nullptr, known_tie);   // ranks >= kmax are never read
    // labels in predicted order (K0) and in ideal order (L; values only, so ties are immaterial)
    if (count_hist) {
        // every label of the row is an integer 0..4 (the caller's fast path checked): level counts by the group, then as below
        int* cnt = reinterpret_cast<int*>(sm.L);                 // (L is written after the counts are read)
        if (grp_tid<WARP>() < 4) cnt[grp_tid<WARP>()] = 0;
        grp_sync<WARP>();
        int c[4] = {0, 0, 0, 0};
        for (int i = grp_tid<WARP>(); i < p.n; i += grp_size<WARP>()) {
            const int lv = (int)yt[i];
#pragma unroll
            for (int L = 1; L <= 4; ++L) c[L - 1] += (lv == L) ? 1 : 0;
        }
#pragma unroll
        for (int L = 0; L < 4; ++L) {
            const int w = __reduce_add_sync(LTR_FULL_MASK, c[L]);
            if ((grp_tid<WARP>() & 31) == 0 && w) atomicAdd(&cnt[L], w);
        }
        grp_sync<WARP>();
        const int c4 = cnt[3], c3 = c4 + cnt[2], c2 = c3 + cnt[1], c1 = c2 + cnt[0];
        grp_sync<WARP>();
        for (int i = grp_tid<WARP>(); i < p.len; i += grp_size<WARP>()) {
            sm.L[i] = (i < p.n) ? (float)((c4 > i) + (c3 > i) + (c2 > i) + (c1 > i)) : ninf;
            if (i < p.n) sm.K0[i] = yt[sm.I[i]];
        }
        grp_sync<WARP>();
    } else if (have_hist) {
        const int c4 = hist >> 24, c3 = c4 + ((hist >> 16) & 255), c2 = c3 + ((hist >> 8) & 255), c1 = c2 + (hist & 255);
        for (int i = grp_tid<WARP>(); i < p.len; i += grp_size<WARP>()) {
            sm.L[i] = (i < p.n) ? (float)((c4 > i) + (c3 > i) + (c2 > i) + (c1 > i)) : ninf;
            if (i < p.n) sm.K0[i] = yt[sm.I[i]];
        }
        grp_sync<WARP>();
    } else {
        for (int i = grp_tid<WARP>(); i < p.len; i += grp_size<WARP>()) {
            sm.L[i] = (i < p.n) ? yt[i] : ninf;
            if (i < p.n) sm.K0[i] = yt[sm.I[i]];
        }
        block_bitonic_desc<WARP>(sm.L, p.len);
    }
    if (grp_tid<WARP>() < p.nk) {
        int k = p.ks[grp_tid<WARP>()];
        if (k <= 0 || k > p.n) k = p.n;
        if (!F64) {
            float got = 0.f, ideal = 0.f;
            for (int r = 0; r < k; ++r) {
                const float d = p.disc32[r];
                got += (exp2f(sm.K0[r]) - 1.0f) / d;
                ideal += (exp2f(sm.L[r]) - 1.0f) / d;
            }
            p.out32[(size_t)b * p.nk + grp_tid<WARP>()] = got / ideal;
        } else {
            double got = 0.0, ideal = 0.0;
            for (int r = 0; r < k; ++r) {
                const double d = p.disc64[r];
                const double lg = (double)sm.K0[r], li = (double)sm.L[r];
                got += (p.exp_gains ? exp2(lg) - 1.0 : lg) / d;
                ideal += (p.exp_gains ? exp2(li) - 1.0 : li) / d;
            }
            p.out64[(size_t)b * p.nk + grp_tid<WARP>()] = p.raw_dcg ? got : ((ideal == 0.0) ? p.empty_value : got / ideal);
        }
    }
    grp_sync<WARP>();
}

// Without flags: every query.  With flags (a top-k fast path ran in front): only the queries q it marked with
// flags[1 + q] != 0.  CTA c owns the contiguous range of ceil(B / grid) flags [c R, (c + 1) R): its threads read them
// coalesced, the marked ones are collected in a shared list (in index order: the list is filled range by range of
// NDCG_MARK_LIST flags) and sorted one after the other.  A CTA whose range holds no mark (every CTA, on scores without
// ties) leaves after one read per thread and one barrier; a heavily marked batch is spread over every CTA.
constexpr int NDCG_MARK_LIST = 4 * RANK_THREADS;        // flags looked at per pass
template <bool F64>
__global__ void __launch_bounds__(RANK_THREADS) ndcg_kernel(const NdcgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int list[NDCG_MARK_LIST];
    __shared__ int nlist;
    RankSmem sm(smem_raw, p.len);
    pdl_prologue();
    if (!p.flags) {
        for (int b = blockIdx.x; b < p.B; b += gridDim.x) ndcg_sorted_query<F64>(p, sm, b);
        return;
    }
    const int per = (p.B + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = (int)blockIdx.x * per, hi = min(p.B, lo + per);
    for (int base = lo; base < hi; base += NDCG_MARK_LIST) {
        int mine[NDCG_MARK_LIST / RANK_THREADS];
        int any = 0;
#pragma unroll
        for (int j = 0; j < NDCG_MARK_LIST / RANK_THREADS; ++j) {
            const int q = base + j * RANK_THREADS + (int)threadIdx.x;
            mine[j] = (q < hi) ? __ldcg(p.flags + 1 + q) : 0;
            any |= mine[j];
        }
        if (!__syncthreads_or(any)) continue;
        if (threadIdx.x == 0) nlist = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NDCG_MARK_LIST / RANK_THREADS; ++j)
            if (mine[j]) list[atomicAdd(&nlist, 1)] = (base + j * RANK_THREADS + (int)threadIdx.x) | ((mine[j] == 2) ? (1 << 30) : 0);
        __syncthreads();
        const int cnt = nlist;
        for (int i = 0; i < cnt; ++i) {                                   // (any order: one output row per query)
            // flag 2: the fast path saw a tie among the first kmax + 1 ranks of a row with integer labels 0..4
            const bool tie_only = (list[i] >> 30) != 0;
            ndcg_sorted_query<F64>(p, sm, list[i] & ~(1 << 30), tie_only, false, 0, tie_only);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Top-k fast path of the NDCG sweep (all k <= 31): no full sort.
//   Every lane sorts its ITEMS (key, label) pairs once with a sorting network in registers; after that the
//   warp's largest remaining key is always at the head of some lane's queue, so one of the kmax rounds is:
//   REDUX.MAX over the heads, ballot of the lanes whose head equals it, shuffle the winner's label to everybody,
//   the winner pops its queue (ITEMS - 1 predicated register moves).  ~10 + 2 ITEMS issue slots per round.
//   Two lanes with the same head, or a popped lane whose next key equals the one it just gave, mean a tie group
//   reaches into the top kmax ranks: the query is appended to the flagged list and sorted in full by
//   ndcg_kernel with the caller's tie rule (std::sort replay included).  A tie group entirely below rank kmax
//   cannot change the top-k ranks.
//   Ideal order: labels that are integers 0..7 (MSLR / Istella: 0..4) are counted per level with ballots and rank r
//   reads its label off the cumulative counts; other labels send the query to the full sort as well.
//   DCG sums: lane r holds the term of rank r; the running sums are formed in rank order (the reference
//   accumulates sequentially in fp32 / fp64) and written where r + 1 equals a k of the sweep.
// Rows of up to 256 docs: one warp per query (ITEMS = 4 or 8).  Longer rows (<= 1024): one CTA of 4 warps per
// query, every warp reduces its 256-doc quarter to kmax candidates, warp 0 merges the 4 kmax candidates.
// 8 n bytes of HBM traffic per query.
__device__ __forceinline__ int topk_key(float f) {
    if (f != f) return 0x7fffffff;                                // NaN first (at::native KeyValueCompDesc)
    return float_order_key(f + 0.0f);                             // -0 -> +0 so that key equality == float equality
}

constexpr int TOPK_TAKEN = (int)0x80000000;

__device__ __forceinline__ void ce_desc(int& ka, float& la, int& kb, float& lb) {
    const bool sw = kb > ka;
    const int k0 = sw ? kb : ka, k1 = sw ? ka : kb;
    const float l0 = sw ? lb : la, l1 = sw ? la : lb;
    ka = k0; kb = k1; la = l0; lb = l1;
}

// descending sorting networks for 4 and 8 register-resident pairs (Batcher odd-even merge sort)
template <int ITEMS>
__device__ __forceinline__ void lane_sort_desc(int (&k)[ITEMS], float (&l)[ITEMS]) {
#define LTR_CE(a, b) ce_desc(k[a], l[a], k[b], l[b])
    if (ITEMS == 4) {
        LTR_CE(0, 1); LTR_CE(2, 3); LTR_CE(0, 2); LTR_CE(1, 3); LTR_CE(1, 2);
    } else {
        LTR_CE(0, 1); LTR_CE(2, 3); LTR_CE(4 % ITEMS, 5 % ITEMS); LTR_CE(6 % ITEMS, 7 % ITEMS);
        LTR_CE(0, 2); LTR_CE(1, 3); LTR_CE(4 % ITEMS, 6 % ITEMS); LTR_CE(5 % ITEMS, 7 % ITEMS);
        LTR_CE(1, 2); LTR_CE(5 % ITEMS, 6 % ITEMS);
        LTR_CE(0, 4 % ITEMS); LTR_CE(1, 5 % ITEMS); LTR_CE(2, 6 % ITEMS); LTR_CE(3, 7 % ITEMS);
        LTR_CE(2, 4 % ITEMS); LTR_CE(3, 5 % ITEMS);
        LTR_CE(1, 2); LTR_CE(3, 4 % ITEMS); LTR_CE(5 % ITEMS, 6 % ITEMS);
    }
#undef LTR_CE
}

// kmax tournament rounds over the lane queues.  Lane r (< kmax) returns the key / label of rank r;
// `tie` is set (warp-uniform) if a tie group reaches into the first kmax ranks.
template <int ITEMS>
__device__ __forceinline__ void topk_rounds(int (&k)[ITEMS], float (&l)[ITEMS], int kmax, int& key_r, float& lab_r,
                                            int& tie) {
    const int lane = threadIdx.x & 31;
    int local_tie = 0;
    unsigned multi = 0;
    key_r = TOPK_TAKEN;
    lab_r = 0.f;
    for (int r = 0; r < kmax; ++r) {
        const int wk = __reduce_max_sync(LTR_FULL_MASK, k[0]);
        const unsigned own = __ballot_sync(LTR_FULL_MASK, k[0] == wk && wk != TOPK_TAKEN);
        multi |= own & (own - 1);                               // more than one lane holds the maximum
        const int owner = own ? __ffs(own) - 1 : 0;
        const float lab = __shfl_sync(LTR_FULL_MASK, l[0], owner);
        if (lane == r) { key_r = wk; lab_r = (wk == TOPK_TAKEN) ? 0.f : lab; }
        const bool me = (lane == owner) && (own != 0);          // the winner pops its queue (predicated, no branch)
#pragma unroll
        for (int i = 0; i + 1 < ITEMS; ++i) k[i] = me ? k[i + 1] : k[i], l[i] = me ? l[i + 1] : l[i];
        k[ITEMS - 1] = me ? TOPK_TAKEN : k[ITEMS - 1];
        local_tie |= (me && k[0] == wk) ? 1 : 0;                // the same key again inside this lane
    }
    tie = (multi != 0) || __any_sync(LTR_FULL_MASK, local_tie);
}

// label at ideal rank `lane` from the per-level counts of integer labels 1..7 (cnt[L] = #docs with label L)
__device__ __forceinline__ float ideal_label_from_counts(const int (&cnt)[8]) {
    const int lane = threadIdx.x & 31;
    int ge = 0, cum = 0;
#pragma unroll
    for (int L = 7; L >= 1; --L) {
        cum += cnt[L];                                          // docs with label >= L
        ge += (cum > lane) ? 1 : 0;
    }
    return (float)ge;                                           // labels fill the ranks from the top level down
}

}  // namespace ltr
#include "ndcg_keys.cuh"
namespace ltr {

template <bool F64>
__device__ __forceinline__ void write_ndcg_sums(const NdcgParams& p, int b, int myk, float lab_pred, float lab_ideal) {
    const int lane = threadIdx.x & 31;
    // labels on this path are integers 0..7: 2^l - 1 is exact integer arithmetic
    const int gp = (1 << (int)lab_pred) - 1, gi = (1 << (int)lab_ideal) - 1;
    if (!F64) {
        const float d = (lane < p.kmax) ? p.disc32[lane] : 1.0f;
        const float tp = (float)gp / d, ti = (float)gi / d;
        float got = 0.f, ideal = 0.f, got_k = 0.f, ideal_k = 0.f;
        for (int r = 0; r < p.kmax; ++r) {
            got += __shfl_sync(LTR_FULL_MASK, tp, r);
            ideal += __shfl_sync(LTR_FULL_MASK, ti, r);
            const bool mine = (myk == r + 1);
            got_k = mine ? got : got_k;
            ideal_k = mine ? ideal : ideal_k;
        }
        if (myk > 0) p.out32[(size_t)b * p.nk + lane] = got_k / ideal_k;
    } else {
        const double d = (lane < p.kmax) ? p.disc64[lane] : 1.0;
        const double tp = (p.exp_gains ? (double)gp : (double)lab_pred) / d;
        const double ti = (p.exp_gains ? (double)gi : (double)lab_ideal) / d;
        double got = 0.0, ideal = 0.0, got_k = 0.0, ideal_k = 0.0;
        for (int r = 0; r < p.kmax; ++r) {
            got += __shfl_sync(LTR_FULL_MASK, tp, r);
            ideal += __shfl_sync(LTR_FULL_MASK, ti, r);
            const bool mine = (myk == r + 1);
            got_k = mine ? got : got_k;
            ideal_k = mine ? ideal : ideal_k;
        }
        if (myk > 0) p.out64[(size_t)b * p.nk + lane] = p.raw_dcg ? got_k : ((ideal_k == 0.0) ? p.empty_value : got_k / ideal_k);
    }
}

// loads ITEMS docs per lane starting at `base`, returns false (warp-uniform) if a label is not an integer 0..7;
// cnt[L] = warp-wide number of docs with label L (L = 1..7)
template <int ITEMS>
__device__ __forceinline__ bool topk_load(const NdcgParams& p, int b, int base, int limit, int (&k)[ITEMS], float (&l)[ITEMS],
                                          int (&cnt)[8]) {
    const int lane = threadIdx.x & 31;
    const float* sp = p.y_pred + (size_t)b * p.n;
    const float* yp = p.y_true + (size_t)b * p.n;
    bool small_int = true;
    int lmax = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int i = base + lane + 32 * j;
        const bool in = i < limit;
        const float s = in ? ld_stream(sp + i) : 0.f;
        l[j] = in ? ld_stream(yp + i) : 0.f;
        k[j] = in ? topk_key(s) : TOPK_TAKEN;
        const int lv = (int)fminf(fmaxf(l[j], -1.f), 8.f);
        small_int = small_int && (!in || ((float)lv == l[j] && lv >= 0 && lv <= 7));
        lmax = max(lmax, in ? lv : 0);
    }
    small_int = __all_sync(LTR_FULL_MASK, small_int);
    lmax = min(__reduce_max_sync(LTR_FULL_MASK, lmax), 7);
#pragma unroll
    for (int L = 1; L <= 7; ++L) {
        int c = 0;
        if (L <= lmax) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
                c += __popc(__ballot_sync(LTR_FULL_MASK, (base + lane + 32 * j < limit) && l[j] == (float)L));
        }
        cnt[L] = c;
    }
    cnt[0] = 0;
    return small_int;
}

template <int ITEMS, bool F64>
__global__ void __launch_bounds__(256) ndcg_topk_kernel(const NdcgParams p) {
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * 8;
    int myk = -1;                                                 // lane j writes the j-th k of the sweep
    for (int j = 0; j < p.nk; ++j) myk = (lane == j) ? p.ks[j] : myk;
    for (int b = gwarp; b < p.B; b += nwarps) {
        int k[ITEMS], cnt[8];
        float l[ITEMS];
        const bool ok = topk_load<ITEMS>(p, b, 0, p.n, k, l, cnt);
        lane_sort_desc<ITEMS>(k, l);
        int key_r, tie;
        float lab_pred;
        topk_rounds<ITEMS>(k, l, p.kmax, key_r, lab_pred, tie);
        if (lane == 0) p.flags[1 + b] = (tie || !ok) ? 1 : 0;     // marked queries are left to the full sort
        if (tie || !ok) continue;
        write_ndcg_sums<F64>(p, b, myk, lab_pred, ideal_label_from_counts(cnt));
    }
}

template <bool F64>
__global__ void __launch_bounds__(128) ndcg_topk_wide_kernel(const NdcgParams p) {
    __shared__ int cand_key[4][32];
    __shared__ float cand_lab[4][32];
    __shared__ int hist[4][8];
    __shared__ int sh_bad[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int myk = -1;
    for (int j = 0; j < p.nk; ++j) myk = (lane == j) ? p.ks[j] : myk;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        int k[8], cnt[8];
        float l[8];
        const bool ok = topk_load<8>(p, b, warp * 256, p.n, k, l, cnt);
        lane_sort_desc<8>(k, l);
        int key_r, tie;
        float lab_r;
        topk_rounds<8>(k, l, p.kmax, key_r, lab_r, tie);
        cand_key[warp][lane] = key_r;                             // TAKEN beyond the quarter's docs / beyond kmax
        cand_lab[warp][lane] = lab_r;
#pragma unroll
        for (int L = 0; L < 8; ++L)
            if (lane == L) hist[warp][L] = cnt[L];
        if (lane == 0) sh_bad[warp] = (tie || !ok) ? 1 : 0;
        __syncthreads();
        if (warp == 0) {
            int mk[4], tot[8];
            float ml[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) mk[q] = cand_key[q][lane], ml[q] = cand_lab[q][lane];
#pragma unroll
            for (int L = 0; L < 8; ++L) tot[L] = (hist[0][L] + hist[1][L]) + (hist[2][L] + hist[3][L]);
            const int bad = (sh_bad[0] | sh_bad[1]) | (sh_bad[2] | sh_bad[3]);
            lane_sort_desc<4>(mk, ml);
            int key_m, tie_m;
            float lab_pred;
            topk_rounds<4>(mk, ml, p.kmax, key_m, lab_pred, tie_m);
            if (lane == 0) p.flags[1 + b] = (bad || tie_m) ? 1 : 0;
            if (!(bad || tie_m)) write_ndcg_sums<F64>(p, b, myk, lab_pred, ideal_label_from_counts(tot));
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
            block_sort_desc<false, RANK_THREADS>(sm.K, sm.I, sm.K0, n, p.len, TIES_STDSORT);
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
            block_sort_desc<false, RANK_THREADS>(sm.K, sm.I, sm.K0, n, p.len, TIES_STDSORT);
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
    const int cap = device_sm_count() * 8;
    bool marked_only = false, keys_done = false;
    if (kmax <= TPQ_K - 1 && p.n % 4 == 0 && p.flags &&
        (reinterpret_cast<uintptr_t>(p.y_true) | reinterpret_cast<uintptr_t>(p.y_pred)) % 16 == 0) {
        // one thread per query (ndcg_keys.cuh / ndcg_stream.cuh), then the full sort for the queries it marked.
        // The opt-in shared-memory size is a per-device function attribute: set on every call (cheap), never cached
        // in a process-wide flag (a second device in one process would otherwise launch without it).
        // warps per tile of 32 queries: long rows are split chunk-wise over the warps of a CTA
        // rows of 2..16 whole chunks with scratch from the caller: the label-in-key kernel chunk by chunk, then a merge
        const bool chunked = p.partial && !f64 && p.n > TPQ_N && p.n % TPQ_N == 0 && p.n <= RANK_MAX_N &&
                             (long long)p.B * (p.n / TPQ_N) < (1ll << 30);
        if (chunked) {
            NdcgParams pv = p;
            pv.B = p.B * (p.n / TPQ_N);
            pv.n = TPQ_N;
            const int vt = (pv.B + 31) / 32, vcap = device_sm_count() * KEY_CTAS;
            LTR_CUDA(cudaFuncSetAttribute(ndcg_key_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KEY_SMEM));
            int rc = launch_pdl(ndcg_key_kernel<false>, vt < vcap ? vt : vcap, TPQ_THREADS, KEY_SMEM, st, pv);
            if (rc) return set_cuda_error(rc, "ltr_ndcg_at_k (chunk keys)");
            int G = 2;                                              // lanes per query in the merge (ndcg_keys.cuh)
            while (G < p.n / TPQ_N && G < 32) G <<= 1;
            const int per_cta = 128 / G;
            rc = launch_pdl(ndcg_chunk_merge_kernel, (p.B + per_cta - 1) / per_cta, 128, 0, st, p);
            if (rc) return set_cuda_error(rc, "ltr_ndcg_at_k (chunk merge)");
        }
        const int share = p.n <= TPQ_N ? 1 : (p.n <= 2 * TPQ_N ? 2 : 4);
        const int tiles = (p.B + 31) / 32;
        const bool keys = share == 1;          // rows of one chunk: label-in-key kernel (ndcg_keys.cuh), which also sorts
                                               // the queries it marks -- no second launch
        const int gcap = device_sm_count() * (keys ? KEY_CTAS : 2);
        int g = (tiles * share + TPQ_WARPS - 1) / TPQ_WARPS;
        g = g < gcap ? g : gcap;
        if (keys) g = tiles < gcap ? tiles : gcap;     // tile t -> CTA t % g: every SM gets its share of a small batch
        int rc = 0;
#define LTR_TPQ(C)                                                                                                      \
    do {                                                                                                                \
        if (f64) {                                                                                                      \
            LTR_CUDA(cudaFuncSetAttribute(ndcg_tpq_wide_kernel<true, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          (int)TPQ_SMEM));                                                              \
            rc = launch_pdl(ndcg_tpq_wide_kernel<true, C>, g, TPQ_THREADS, TPQ_SMEM, st, p);                            \
        } else {                                                                                                        \
            LTR_CUDA(cudaFuncSetAttribute(ndcg_tpq_wide_kernel<false, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)TPQ_SMEM));                                                              \
            rc = launch_pdl(ndcg_tpq_wide_kernel<false, C>, g, TPQ_THREADS, TPQ_SMEM, st, p);                           \
        }                                                                                                               \
    } while (0)
        if (chunked) {
            // (launched above)
        } else if (keys) {
            if (f64) {
                LTR_CUDA(cudaFuncSetAttribute(ndcg_key_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KEY_SMEM));
                rc = launch_pdl(ndcg_key_kernel<true>, g, TPQ_THREADS, KEY_SMEM, st, p);
            } else {
                LTR_CUDA(cudaFuncSetAttribute(ndcg_key_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KEY_SMEM));
                rc = launch_pdl(ndcg_key_kernel<false>, g, TPQ_THREADS, KEY_SMEM, st, p);
            }
        } else if (share == 2) { LTR_TPQ(2); } else { LTR_TPQ(4); }
#undef LTR_TPQ
        if (rc) return set_cuda_error(rc, "ltr_ndcg_at_k (one thread per query)");
        marked_only = true;
        keys_done = keys && !chunked;
    } else if (kmax <= 31 && p.n <= 1024 && p.flags) {
        // top-k fast path (one warp per query), then the full sort for the queries it marked
        int g = (p.B + 7) / 8;
        g = g < cap ? g : cap;
#define LTR_TOPK(ITEMS)                                                      \
        if (f64) ndcg_topk_kernel<ITEMS, true><<<g, 256, 0, st>>>(p);            \
        else ndcg_topk_kernel<ITEMS, false><<<g, 256, 0, st>>>(p)
        if (p.n <= 128) { LTR_TOPK(4); }
        else if (p.n <= 256) { LTR_TOPK(8); }
        else {
            const int gw = p.B < cap ? p.B : cap;          // ties of any rule go to the full sort there
            if (f64) ndcg_topk_wide_kernel<true><<<gw, 128, 0, st>>>(p);
            else ndcg_topk_wide_kernel<false><<<gw, 128, 0, st>>>(p);
        }
#undef LTR_TOPK
        marked_only = true;
    } else {
        p.flags = nullptr;
    }
    int gs = rank_grid(p.B);
    if (marked_only) {
        const int ranges = (p.B + 7) / 8;              // a small, heavily marked batch still spreads over many CTAs
        gs = ranges < cap ? ranges : cap;
    }
#ifdef LTR_NDCG_SKIP_MARKED
    if (marked_only) return 0;                 // timing experiments only: the marked queries are left unsorted
#endif
    if (keys_done) return 0;                   // the label-in-key kernel sorts the queries it marks itself
    if (f64) {
        if (int rc = rank_prepare(ndcg_kernel<true>, smem)) return rc;
        const int rc = launch_pdl(ndcg_kernel<true>, gs, RANK_THREADS, smem, st, p);
        return rc ? set_cuda_error(rc, "ltr_ndcg_at_k") : 0;
    } else {
        if (int rc = rank_prepare(ndcg_kernel<false>, smem)) return rc;
        const int rc = launch_pdl(ndcg_kernel<false>, gs, RANK_THREADS, smem, st, p);
        return rc ? set_cuda_error(rc, "ltr_ndcg_at_k") : 0;
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

LTR_API size_t ltr_ndcg_at_k_scratch_bytes(int Q, int n) {
    using namespace ltr;
    if (Q <= 0 || n <= TPQ_N || n % TPQ_N != 0 || n > RANK_MAX_N || (long long)Q * (n / TPQ_N) >= (1ll << 30)) return 0;
    return (size_t)13 * sizeof(float) * (size_t)Q * (size_t)(n / TPQ_N);
}

LTR_API int ltr_ndcg_at_k_ws(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                             const float* disc32, float* out, int* flags, void* scratch, size_t scratch_bytes, void* stream) {
    ltr::NdcgParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.B = Q; p.n = n; p.tie_rule = tie_rule; p.disc32 = disc32; p.out32 = out;
    p.flags = flags;
    if (!disc32 || !out) return ltr::set_error(LTR_EINVAL, "ltr_ndcg_at_k_ws: null table / output");
    const size_t need = ltr_ndcg_at_k_scratch_bytes(Q, n);
    if (scratch && need && scratch_bytes < need) return ltr::set_error(LTR_EINVAL, "ltr_ndcg_at_k_ws: scratch too small");
    if (scratch && need && (reinterpret_cast<uintptr_t>(scratch) & 15)) return ltr::set_error(LTR_EINVAL, "ltr_ndcg_at_k_ws: scratch not 16-byte aligned");
    p.partial = (scratch && need) ? static_cast<float*>(scratch) : nullptr;
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

// NDCG@k sweep from HOST buffers (pinned for asynchrony): H2D of both [Q, n] arrays, the sweep kernel, D2H of [Q, nk].
LTR_API size_t ltr_ndcg_at_k_host_scratch_bytes(int Q, int n, int nk) {
    if (Q <= 0 || n <= 0 || nk <= 0) return 0;
    const size_t qn = (((size_t)Q * n + 3) & ~(size_t)3);
    const size_t qk = (((size_t)Q * nk + 3) & ~(size_t)3);
    // (+ the per-chunk top lists of long rows: ltr_ndcg_at_k_ws)
    return (2 * qn + qk + (((size_t)Q + 1 + 3) & ~(size_t)3)) * sizeof(float) + ltr_ndcg_at_k_scratch_bytes(Q, n);
}

LTR_API int ltr_ndcg_at_k_host(const float* h_y_true, const float* h_y_pred, int Q, int n, const int* ks, int nk,
                               int tie_rule, const float* disc32, float* h_out, void* d_scratch, size_t scratch_bytes,
                               void* stream) {
    using namespace ltr;
    if (!h_y_true || !h_y_pred || !h_out || !d_scratch || Q <= 0 || n <= 0 || nk <= 0)
        return set_error(LTR_EINVAL, "ltr_ndcg_at_k_host: null pointer or non-positive size");
    if (scratch_bytes < ltr_ndcg_at_k_host_scratch_bytes(Q, n, nk))
        return set_error(LTR_EINVAL, "ltr_ndcg_at_k_host: scratch too small");
    const size_t qn = (((size_t)Q * n + 3) & ~(size_t)3);
    const size_t qk = (((size_t)Q * nk + 3) & ~(size_t)3);
    float* d_true = reinterpret_cast<float*>(d_scratch);
    float* d_pred = d_true + qn;
    float* d_out = d_pred + qn;
    int* d_flags = reinterpret_cast<int*>(d_out + qk);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)Q * n * sizeof(float);
    LTR_CUDA(cudaMemcpyAsync(d_true, h_y_true, bytes, cudaMemcpyHostToDevice, st));
    LTR_CUDA(cudaMemcpyAsync(d_pred, h_y_pred, bytes, cudaMemcpyHostToDevice, st));
    void* d_chunks = reinterpret_cast<float*>(d_flags) + (((size_t)Q + 1 + 3) & ~(size_t)3);
    if (int rc = ltr_ndcg_at_k_ws(d_true, d_pred, Q, n, ks, nk, tie_rule, disc32, d_out, d_flags, d_chunks,
                                  ltr_ndcg_at_k_scratch_bytes(Q, n), stream))
        return rc;
    LTR_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)Q * nk * sizeof(float), cudaMemcpyDeviceToHost, st));
    return 0;
}

LTR_API int ltr_dcg_at_k_f64(const float* y_true, const float* y_pred, int Q, int n, const int* ks, int nk, int tie_rule,
                             int exponential_gains, const double* disc64, double* out, int* flags, void* stream) {
    ltr::NdcgParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.B = Q; p.n = n; p.tie_rule = tie_rule; p.disc64 = disc64; p.out64 = out;
    p.flags = flags;
    p.exp_gains = exponential_gains; p.empty_value = 0.0; p.raw_dcg = 1;
    if (!disc64 || !out) return ltr::set_error(LTR_EINVAL, "ltr_dcg_at_k_f64: null table / output");
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
