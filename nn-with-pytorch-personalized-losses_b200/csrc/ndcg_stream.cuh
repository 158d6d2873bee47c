// K4 fast path, second generation: NDCG@k sweep with ONE THREAD PER QUERY (replaces utils/metrics.py:48-104 for
// sweeps whose largest k is <= 10, n % 4 == 0; rows longer than 128 docs are processed in 128-doc chunks).
//
// The warp-per-query top-k kernel spends ~1,000 issue slots per query on cross-lane tournaments.  Here a warp
// owns a tile of 32 queries and works in two phases:
//
//   phase 1 (lane = 4 consecutive docs of one row, coalesced 128-bit loads, next batch always in flight): the
//     row's scores go to shared memory in a layout the second phase reads without bank conflicts, the labels
//     packed to 4 bits; the label histogram of the row (levels 1..4, one REDUX) and its "not on the fast
//     path" flag (NaN score, label not an integer 0..4) stay in the lane of that row;
//   phase 2 (lane = one query, no cross-lane traffic at all): the 11th largest of the 32 quad maxima is a lower
//     bound L of the 11th largest score (a pruned Batcher network in registers); the docs with score >= L
//     (13.5 on average) are the only candidates: they are inserted, with their labels, into a sorted top-11 list
//     held in registers; DCG and ideal DCG (from the histogram) are accumulated in rank order exactly as the
//     reference does (sequential fp32 / fp64 sums).
//
// Exact-tie handling is unchanged: a query with equal scores among its first kmax + 1 ranks (whose order depends
// on the reference's sort) is flagged, and ndcg_kernel sorts it in full with the caller's tie rule.
// HBM traffic: 8 n bytes per query + the outputs, read once.
#pragma once
#include "ltr_common.cuh"

namespace ltr {

constexpr int TPQ_K = 11;              // ranks kept per query: kmax <= 10 plus one for the boundary tie check
constexpr int TPQ_WARPS = 4;           // independent warps per CTA
constexpr int TPQ_THREADS = 32 * TPQ_WARPS;
constexpr int TPQ_CTAS = 3;            // CTAs per SM (shared memory: 3 x 4 x 19,072 B)
constexpr int TPQ_N = 128;             // docs per row (one chunk)
constexpr int TPQ_BATCH = 8;           // label rows per load batch (one 128-bit load per row and lane in flight)
constexpr int TPQ_S_STRIDE = TPQ_N + 4;    // floats; 132 q + 4 j: conflict-free 128-bit reads per quarter warp
constexpr int TPQ_L_STRIDE = 34;           // 16-bit units; labels of a quad packed to 4 x 4 bits
constexpr int TPQ_Q_STRIDE = 36;           // quad maxima per row: conflict-free 128-bit reads per quarter warp

struct TpqTile {
    float s[32 * TPQ_S_STRIDE];
    unsigned short lab[32 * TPQ_L_STRIDE];
};
constexpr size_t TPQ_SMEM = sizeof(TpqTile) * TPQ_WARPS;

__device__ __forceinline__ float max_nan(float a, float b) {       // NaN if either operand is NaN
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned shr_clamp(unsigned v, unsigned s) {   // v >> s, 0 for s >= 32
    unsigned r;
    asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

// insertion with a payload; an equal key goes BELOW the entries already there (ties are flagged anyway).
// Parallel form: up[j] = x > t[j] is monotone in j (t is sorted), so every slot picks its new value from
// {t[j-1], x, t[j]} independently -- the same 5 K instructions as a bubble chain, but dependency depth 3.
template <int K>
__device__ __forceinline__ void tpq_insert(float (&t)[K], int (&l)[K], float x, int lab) {
    bool up[K];
#pragma unroll
    for (int j = 0; j < K; ++j) up[j] = x > t[j];
#pragma unroll
    for (int j = K - 1; j >= 1; --j) {
        t[j] = up[j - 1] ? t[j - 1] : (up[j] ? x : t[j]);
        l[j] = up[j - 1] ? l[j - 1] : (up[j] ? lab : l[j]);
    }
    t[0] = up[0] ? x : t[0];
    l[0] = up[0] ? lab : l[0];
}

// the 11th largest of 32 values: Batcher's odd-even merge sort pruned to the comparators that output depends on
__device__ __forceinline__ float tpq_select11(float (&v)[32]) {
#define TPQ_CE(a, b)                          \
    {                                         \
        const float hi_ = fmaxf(v[a], v[b]);  \
        v[b] = fminf(v[a], v[b]);             \
        v[a] = hi_;                           \
    }
#include "ndcg_select32.inc"
#undef TPQ_CE
    return v[TPQ_K - 1];
}

// Rows longer than one 128-doc chunk (same two phases per chunk; the top-11 list and the label counts of a query are
// carried from chunk to chunk).  C = warps that share one tile of 32 queries: the chunks of its rows are dealt chunk by chunk to the C
// warps (C = 1: four independent tiles per CTA; C = 4: one tile per CTA, 8 chunks of a 1024-doc row = 2 per warp);
// the C partial top-11 lists and label histograms of a query meet in shared memory and lane q of the first warp
// merges them.
template <bool F64, int C>
__global__ void __launch_bounds__(TPQ_THREADS, C == 1 ? TPQ_CTAS : 2) ndcg_tpq_wide_kernel(const NdcgParams p) {
    extern __shared__ __align__(16) unsigned char tpq_smem[];
    // term[r][l] = gain(l) / discount(r) for ranks r < 10 and labels l <= 4, divided exactly as the reference divides
    __shared__ double term64[F64 ? (TPQ_K - 1) * 5 : 1];
    __shared__ float term32[F64 ? 1 : (TPQ_K - 1) * 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TpqTile* tiles = reinterpret_cast<TpqTile*>(tpq_smem);
    TpqTile& tile = tiles[warp];
    constexpr int TILES_PER_CTA = TPQ_WARPS / C;
    const int sub = warp % C, slot = warp / C;
    const int ntiles = (p.B + 31) >> 5;
    const int ncta = (ntiles + TILES_PER_CTA - 1) / TILES_PER_CTA;
    const int n = p.n;
    const int nchunks = (C == 1) ? 1 : (n + TPQ_N - 1) / TPQ_N;   // C == 1 is launched for rows of one chunk only
    const float ninf = __int_as_float(0xff800000);
    unsigned kends = 0;                                    // bit r: some k of the sweep equals r + 1
    for (int j = 0; j < p.nk; ++j) kends |= 1u << (p.ks[j] - 1);
    pdl_prologue();
    if ((int)threadIdx.x < (TPQ_K - 1) * 5) {
        const int r = (int)threadIdx.x / 5, l = (int)threadIdx.x % 5;
        if (r < p.kmax) {
            if (F64) term64[threadIdx.x] = (p.exp_gains ? (double)((1 << l) - 1) : (double)l) / p.disc64[r];
            else term32[threadIdx.x] = (float)((1 << l) - 1) / p.disc32[r];
        }
    }
    __syncthreads();

    // labels go through registers in batches of TPQ_BATCH rows, the next batch (of the next chunk / tile during
    // phase 2) always in flight; scores go straight to shared memory with 16-byte asynchronous copies, a chunk at once
    float4 y4[TPQ_BATCH];
    auto issue_labels = [&](int qbase, int d0) {            // rows qbase.., docs d0 + 4 lane ..
        const bool in = d0 + 4 * lane < n && d0 + 4 * lane < d0 + TPQ_N;
        const float* yp = p.y_true + (size_t)qbase * n + d0 + 4 * lane;
#pragma unroll
        for (int i = 0; i < TPQ_BATCH; ++i) {
            y4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (qbase + i < p.B && in) y4[i] = ld_stream4(yp + (size_t)i * n);
        }
    };
    {
        const int t0 = (int)blockIdx.x * TILES_PER_CTA + slot;
        if (t0 < ntiles && sub < nchunks) issue_labels(t0 << 5, sub * TPQ_N);
    }

    for (int ct = (int)blockIdx.x; ct < ncta; ct += (int)gridDim.x) {
        const int t = ct * TILES_PER_CTA + slot;
        const int q0 = t << 5;
        const int q = q0 + lane;
        float top[TPQ_K];
        int lab[TPQ_K];
        if (C > 1 || t >= ntiles) {                          // (one chunk per row: initialised where phase 2 starts)
#pragma unroll
            for (int j = 0; j < TPQ_K; ++j) top[j] = ninf, lab[j] = 0;
        }
        int c1 = 0, c2 = 0, c3 = 0, c4 = 0;                  // #docs of the query with label >= 1 / 2 / 3 / 4 ... as counts per level
        int bad = 0;

        if (t < ntiles) {
            for (int c = sub; c < nchunks; c += C) {
                const int d0 = c * TPQ_N;
                const int nc = min(TPQ_N, n - d0);           // docs of this chunk (a multiple of 4)
                const bool lane_in = 4 * lane < nc;
                // ---------------- phase 1: rows -> shared memory, per-row histogram and flag ----------------
                {
                    const float* sp = p.y_pred + (size_t)q0 * n + d0 + 4 * lane;
                    const unsigned dst = smem_u32(&tile.s[4 * lane]);
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        if (q0 + r < p.B && lane_in)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * TPQ_S_STRIDE * 4),
                                         "l"(sp + (size_t)r * n)
                                         : "memory");
                        else
                            *reinterpret_cast<float4*>(&tile.s[r * TPQ_S_STRIDE + 4 * lane]) = make_float4(ninf, ninf, ninf, ninf);
                    }
                }
                unsigned myhist = 0;
                int mybad = 0;
#pragma unroll 1
                for (int rb = 0; rb < 32; rb += TPQ_BATCH) {
                    float4 yc[TPQ_BATCH];
#pragma unroll
                    for (int i = 0; i < TPQ_BATCH; ++i) yc[i] = y4[i];
                    // next batch: further rows of this chunk, else this warp's next chunk, else its next tile
                    if (rb < 32 - TPQ_BATCH) {
                        issue_labels(q0 + rb + TPQ_BATCH, d0);
                    } else if (c + C < nchunks) {
                        issue_labels(q0, d0 + C * TPQ_N);
                    } else {
                        const int tn = (ct + (int)gridDim.x) * TILES_PER_CTA + slot;
                        issue_labels((ct + (int)gridDim.x < ncta && tn < ntiles && sub < nchunks) ? (tn << 5) : p.B, sub * TPQ_N);
                    }
                    unsigned badbits = 0;
#pragma unroll
                    for (int i = 0; i < TPQ_BATCH; ++i) {
                        const int r = rb + i;
                        const float ys[4] = {yc[i].x, yc[i].y, yc[i].z, yc[i].w};
                        unsigned hist = 0, word = 0;
                        bool ok = true;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const unsigned v = __float2uint_rz(ys[cc]);                // negative / NaN -> 0, huge -> 2^32 - 1
                            ok = ok && ((float)v == ys[cc]) && (v <= 4u);
                            hist += shr_clamp(0x01000000u, 32u - 8u * v);              // level L counts in byte L - 1; level 0: nothing
                            word |= (v & 0xfu) << (4 * cc);
                        }
                        tile.lab[r * TPQ_L_STRIDE + lane] = (unsigned short)word;
                        badbits |= ok ? 0u : (1u << i);
                        const unsigned rowhist = __reduce_add_sync(LTR_FULL_MASK, hist);
                        myhist = (lane == r) ? rowhist : myhist;
                    }
                    badbits = __reduce_or_sync(LTR_FULL_MASK, badbits);      // one vote per batch
                    if (lane / TPQ_BATCH == rb / TPQ_BATCH) mybad = (int)((badbits >> (lane % TPQ_BATCH)) & 1u);
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();

                // ---------------- phase 2: one query per lane --------------------------------------------------
                const float* srow = &tile.s[lane * TPQ_S_STRIDE];
                if (C == 1) {
#pragma unroll
                    for (int j = 0; j < TPQ_K; ++j) top[j] = ninf, lab[j] = 0;
                }
                // lower bound of the TPQ_K-th largest score: the TPQ_K-th largest of the 32 quad maxima of this chunk,
                // or the TPQ_K-th best of the chunks before it
                float qm[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(srow + 4 * j);
                    qm[j] = max_nan(max_nan(x.x, x.y), max_nan(x.z, x.w));       // a NaN score sends the query to the full sort
                }
                bool anynan = false;
#pragma unroll
                for (int j = 0; j < 32; ++j) anynan = anynan || (qm[j] != qm[j]);
                const float L = fmaxf(tpq_select11(qm), top[TPQ_K - 1]);
                asm volatile("" ::: "memory");          // read the row again below rather than keeping 128 scores in registers
                // candidate docs: score >= L
                unsigned cand[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(srow + 4 * j);
                    unsigned& w = cand[j >> 3];
                    const int b0 = 4 * (j & 7);
                    w |= (x.x >= L) ? (1u << b0) : 0u;
                    w |= (x.y >= L) ? (2u << b0) : 0u;
                    w |= (x.z >= L) ? (4u << b0) : 0u;
                    w |= (x.w >= L) ? (8u << b0) : 0u;
                }
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int left = nc - 64 * w;                                // docs of the chunk in this 64-doc word
                    const unsigned long long nmask = left >= 64 ? ~0ull : (left <= 0 ? 0ull : ((1ull << left) - 1ull));
                    unsigned long long m = (((unsigned long long)cand[2 * w + 1] << 32) | cand[2 * w]) & nmask;
                    while (m) {
                        const int d = 64 * w + (__ffsll((long long)m) - 1);
                        m &= m - 1;
                        const float x = srow[d] + 0.0f;                          // -0 -> +0
                        const unsigned word = tile.lab[lane * TPQ_L_STRIDE + (d >> 2)];
                        tpq_insert<TPQ_K>(top, lab, x, (int)((word >> (4 * (d & 3))) & 0xfu));
                    }
                }
                bad |= mybad | (anynan ? 1 : 0);
                c4 += (int)(myhist >> 24);
                c3 += (int)((myhist >> 16) & 255u);
                c2 += (int)((myhist >> 8) & 255u);
                c1 += (int)(myhist & 255u);
                __syncwarp();                                    // the tile is free for the next phase 1
            }
        }

        // ---------------- the C partial results of a query meet in the first warp of the tile ----------------
        if (C > 1) {
            constexpr int WORDS = 2 * TPQ_K + 5;
            float* mine = tile.s;                                // [WORDS][32], free after the last phase 2
            if (sub != 0) {
#pragma unroll
                for (int j = 0; j < TPQ_K; ++j) {
                    mine[j * 32 + lane] = top[j];
                    mine[(TPQ_K + j) * 32 + lane] = __int_as_float(lab[j]);
                }
                mine[(2 * TPQ_K + 0) * 32 + lane] = __int_as_float(c1);
                mine[(2 * TPQ_K + 1) * 32 + lane] = __int_as_float(c2);
                mine[(2 * TPQ_K + 2) * 32 + lane] = __int_as_float(c3);
                mine[(2 * TPQ_K + 3) * 32 + lane] = __int_as_float(c4);
                mine[(2 * TPQ_K + 4) * 32 + lane] = __int_as_float(bad);
            }
            __syncthreads();
            if (sub == 0) {
#pragma unroll 1
                for (int o = 1; o < C; ++o) {
                    const float* other = tiles[warp + o].s;
#pragma unroll
                    for (int j = 0; j < TPQ_K; ++j)
                        tpq_insert<TPQ_K>(top, lab, other[j * 32 + lane], __float_as_int(other[(TPQ_K + j) * 32 + lane]));
                    c1 += __float_as_int(other[(2 * TPQ_K + 0) * 32 + lane]);
                    c2 += __float_as_int(other[(2 * TPQ_K + 1) * 32 + lane]);
                    c3 += __float_as_int(other[(2 * TPQ_K + 2) * 32 + lane]);
                    c4 += __float_as_int(other[(2 * TPQ_K + 3) * 32 + lane]);
                    bad |= __float_as_int(other[(2 * TPQ_K + 4) * 32 + lane]);
                }
            }
            __syncthreads();
            static_assert(WORDS * 32 <= 32 * TPQ_S_STRIDE, "merge records fit the score tile");
        }

        // a tie group that reaches into the first kmax ranks needs the reference's sort order
        // flag 1: labels / scores the fast path does not handle; flag 2: only the tie (labels are integers 0..4): the full
        // sort then skips its sorting network (the order is replayed anyway) and counts the ideal order off a histogram
        int tie_top = 0;
#pragma unroll
        for (int j = 0; j + 1 < TPQ_K; ++j)
            tie_top |= (j < p.kmax && j + 1 < n && top[j] == top[j + 1]) ? 1 : 0;
        if (sub == 0 && t < ntiles && q < p.B) {
            p.flags[1 + q] = bad ? 1 : (tie_top ? 2 : 0);
            bad |= tie_top;
            if (!bad) {
                // ideal order from the histogram: level L fills the ranks below #docs with label >= L
                c3 += c4;
                c2 += c3;
                c1 += c2;
                if (!F64) {
                    float got = 0.f, ideal = 0.f;
                    float* out = p.out32 + (size_t)q * p.nk;
#pragma unroll
                    for (int r = 0; r < TPQ_K - 1; ++r) {
                        if (r < p.kmax) {
                            const int li = (c4 > r) + (c3 > r) + (c2 > r) + (c1 > r);
                            got += term32[5 * r + lab[r]];
                            ideal += term32[5 * r + li];
                            if ((kends >> r) & 1u) {
                                const float val = got / ideal;
                                for (int j = 0; j < p.nk; ++j)
                                    if (p.ks[j] == r + 1) out[j] = val;
                            }
                        }
                    }
                } else {
                    double got = 0.0, ideal = 0.0;
                    double* out = p.out64 + (size_t)q * p.nk;
#pragma unroll
                    for (int r = 0; r < TPQ_K - 1; ++r) {
                        if (r < p.kmax) {
                            const int li = (c4 > r) + (c3 > r) + (c2 > r) + (c1 > r);
                            got += term64[5 * r + lab[r]];
                            ideal += term64[5 * r + li];
                            if ((kends >> r) & 1u) {
                                const double val = p.raw_dcg ? got : ((ideal == 0.0) ? p.empty_value : got / ideal);
                                for (int j = 0; j < p.nk; ++j)
                                    if (p.ks[j] == r + 1) out[j] = val;
                            }
                        }
                    }
                }
            }
        }
    }
}

}  // namespace ltr
