// K4 fast path, fourth generation: NDCG@k sweep for rows of one chunk (n <= 128, n % 4 == 0, largest k <= 10) with the
// LABEL CARRIED INSIDE THE SORT KEY (replaces utils/metrics.py:48-104 on that shape).
//
// The earlier one-thread-per-query kernels (ndcg_stream.cuh) spend two thirds of their instructions moving a payload
// (quad index / label) through the selection network and the insertions.  Here a doc is ONE 32-bit word:
//
//     key = (score bits & ~7) | label,      label in 0..4, score canonicalised (-0 -> +0), compared as fp32
//
// Two keys whose upper 29 bits differ compare exactly like the scores.  Two keys whose upper 29 bits agree (an exact
// tie or a difference in the last 3 mantissa bits: ~1e-4 of all queries on random scores) are detected among the
// first kmax + 1 ranks, and that query alone is redone by its thread with exact compares on the row read back from
// L2; only an exact tie is left to the full sort (ndcg_kernel), as before.
//
//   phase 1 (lane = 4 consecutive docs of one row; KEY_BATCH rows in flight per warp, reissued slot by slot):
//     builds the 4 keys, sorts them in registers (5 compare-exchanges) and writes the quad's largest, second and
//     third key to three shared arrays A, B, C; label validation (integer 0..4), the +-inf / NaN test and the label
//     histogram of the row (one REDUX) stay in the lane of that row.  Validation runs on the FMA pipe
//     (magic-number round trip, squared residues) to balance it against the ALU pipe that does the min/max work.
//   phase 2 (lane = one query): a key-only Batcher network pruned to its first 11 outputs sorts the 32 A keys
//     (2 FMNMX per comparator); L = the 11th is a lower bound of the 11th largest key of the row, so the row's top 11
//     are those 11 A keys plus the few B / C (/ D, rebuilt from L2: 0.3 % of queries) keys above L, which are
//     inserted with 2 FMNMX per slot.  DCG and ideal DCG (from the histogram) are accumulated in rank order exactly
//     as the reference does.
//
// HBM traffic: 8 n bytes per query + the outputs, read once.  Rows with a +-inf score take the full sort too (their
// keys would be NaN patterns).
#pragma once
#include "ndcg_stream.cuh"

namespace ltr {

#ifndef LTR_KEY_BATCH
#define LTR_KEY_BATCH 4
#endif
#ifndef LTR_KEY_CTAS
#define LTR_KEY_CTAS 4
#endif
constexpr int KEY_BATCH = LTR_KEY_BATCH;   // rows in flight per warp (two 128-bit loads per row and lane)
constexpr int KEY_CTAS = LTR_KEY_CTAS;     // CTAs of 4 independent warps per SM
constexpr int KEY_AB_STRIDE = 36;          // floats per row of A / B: conflict-free 128-bit reads per quarter warp
constexpr int KEY_C_STRIDE = 33;           // floats per row of C: read one word at a time
constexpr int KEY_LUT = 625;               // label histogram of a quad by its base-5 code
#ifndef LTR_KEY_AHEAD
#define LTR_KEY_AHEAD 8
#endif
constexpr int KEY_AHEAD = LTR_KEY_AHEAD;   // rows between the L2 prefetch and the demand loads of a warp (a multiple of KEY_BATCH)

struct KeyTile {
    float a[32 * KEY_AB_STRIDE];
    float b[32 * KEY_AB_STRIDE];
    float c[32 * KEY_C_STRIDE];
};
constexpr size_t KEY_LUT_BYTES = (KEY_LUT * 4 + 15) / 16 * 16;
constexpr size_t KEY_SMEM = KEY_LUT_BYTES + sizeof(KeyTile) * TPQ_WARPS;
static_assert(sizeof(float) * 32 * KEY_AB_STRIDE >= (4 * 128 + 3 * 48) * sizeof(float), "a marked row is sorted inside the A array");
static_assert(KEY_LUT_BYTES + sizeof(KeyTile) >= 4096, "a 10-bit quad code read stays inside the CTA's shared memory");

// HBM -> L2 only: the demand loads that follow find the rows in L2, so the few rows a warp can keep in flight in its
// registers cover the L2 latency instead of the DRAM latency.  bytes: a multiple of 16, p: 16-byte aligned.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// (a & ~7) | (b & 7) in one LOP3
__device__ __forceinline__ unsigned merge_low3(unsigned a, unsigned b) {
    unsigned r;
    asm("lop3.b32 %0, %1, 7, %2, 0xB8;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ float key_of(float s, float t) {        // t = label + 2^23: the label sits in its low bits
    return __uint_as_float(merge_low3(__float_as_uint(s + 0.0f), __float_as_uint(t)));
}
__device__ __forceinline__ bool key_near(float a, float b) {       // same upper 29 bits
    return ((__float_as_uint(a) ^ __float_as_uint(b)) < 8u);
}
// v sorted descending; x joins it (the smallest entry falls out).  New v[j] = clamp(x, v[j], old v[j-1]).
template <int K>
__device__ __forceinline__ void key_insert(float (&v)[K], float x) {
#pragma unroll
    for (int j = K - 1; j >= 1; --j) v[j] = fmaxf(v[j], fminf(v[j - 1], x));
    v[0] = fmaxf(v[0], x);
}

// the full sort of a marked query by one warp, out of line: it is rare, and the hot loops stay compact
template <bool F64>
__device__ __noinline__ void key_full_sort(const NdcgParams& p, float* scratch, int b, bool known_tie, bool have_hist,
                                           unsigned hist) {
    RankSmem rs(scratch, p.len);
    ndcg_sorted_query<F64, true>(p, rs, b, known_tie, have_hist, hist);
}

template <bool F64>
__global__ void __launch_bounds__(TPQ_THREADS, KEY_CTAS) ndcg_key_kernel(const __grid_constant__ NdcgParams p) {
    extern __shared__ __align__(16) unsigned char tpq_smem[];
    // term[r][l] = gain(l) / discount(r) for ranks r < 10 and labels l <= 4, divided exactly as the reference divides
    __shared__ double term64[F64 ? (TPQ_K - 1) * 8 : 1];
    __shared__ float term32[F64 ? 1 : (TPQ_K - 1) * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* lut = reinterpret_cast<unsigned*>(tpq_smem);
    KeyTile& tile = reinterpret_cast<KeyTile*>(tpq_smem + KEY_LUT_BYTES)[warp];
    // tile t goes to CTA t % grid, warp t / grid: a batch of fewer tiles than warps still loads every SM alike
    const int gwarp = warp * (int)gridDim.x + (int)blockIdx.x;
    const int nwarps = (int)gridDim.x * TPQ_WARPS;
    const int ntiles = (p.B + 31) >> 5;
    const int n = p.n;
    const float SENT = __uint_as_float(0xff7ffff8u);       // docs beyond n: below every finite score, label 0
    const float MAGIC = 8388608.0f;                        // 2^23: y + MAGIC has the integer y in its low mantissa bits
    const float finf = __int_as_float(0x7f800000);
    const bool lane_in = 4 * lane < n;                     // this lane's quad exists (n % 4 == 0)
    unsigned kends = 0;                                    // bit r: some k of the sweep equals r + 1
    for (int j = 0; j < p.nk; ++j) kends |= 1u << (p.ks[j] - 1);
    // Before the programmatic dependency is resolved (the previous kernel on the stream may still run): nothing is read
    // or written yet -- the histogram table is pure arithmetic, and the first rows only move from HBM to L2, where a late
    // write of the previous kernel would still be seen by the loads below.
    asm volatile("griddepcontrol.launch_dependents;");
    // quad code c = l0 + 5 l1 + 25 l2 + 125 l3 -> its label histogram: level L counts in byte L - 1; level 0: nothing
    if ((int)threadIdx.x < 125) {
        const int l0 = (int)threadIdx.x % 5, l1 = ((int)threadIdx.x / 5) % 5, l2 = (int)threadIdx.x / 25;
        const unsigned h = (l0 ? 1u << (8 * (l0 - 1)) : 0u) + (l1 ? 1u << (8 * (l1 - 1)) : 0u) + (l2 ? 1u << (8 * (l2 - 1)) : 0u);
        lut[threadIdx.x] = h;
#pragma unroll
        for (int l3 = 1; l3 < 5; ++l3) lut[threadIdx.x + 125 * l3] = h + (1u << (8 * (l3 - 1)));
    }
    // rows [q, q + KEY_BATCH) of both arrays -> L2 (one lane, two instructions)
    auto prefetch_rows = [&](int q) {
        const int rows = min(KEY_BATCH, p.B - q);
        if (KEY_AHEAD > 0 && lane == 0 && rows > 0) {
            prefetch_l2_bulk(p.y_pred + (size_t)q * n, (unsigned)(rows * n * 4));
            prefetch_l2_bulk(p.y_true + (size_t)q * n, (unsigned)(rows * n * 4));
        }
    };
    if (gwarp < ntiles) {
        const int rows = min(KEY_AHEAD + KEY_BATCH, p.B - (gwarp << 5));
        if (KEY_AHEAD > 0 && lane == 0) {
            prefetch_l2_bulk(p.y_pred + (size_t)(gwarp << 5) * n, (unsigned)(rows * n * 4));
            prefetch_l2_bulk(p.y_true + (size_t)(gwarp << 5) * n, (unsigned)(rows * n * 4));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // KEY_BATCH rows in flight, reissued slot by slot.  Lanes beyond n keep (SENT, 0); rows beyond B keep the
    // values of an earlier row (their queries are never written).
    float4 s4[KEY_BATCH], y4[KEY_BATCH];
#pragma unroll
    for (int i = 0; i < KEY_BATCH; ++i) s4[i] = make_float4(SENT, SENT, SENT, SENT), y4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue_row = [&](float4& s, float4& y, const float* sb, const float* yb, int q, unsigned off) {
        if (q < p.B && lane_in) {
            s = ld_stream4(sb + off);
            y = ld_stream4(yb + off);
        }
    };
    const float* sbase = p.y_pred + (size_t)min(gwarp, ntiles - 1) * 32 * n + 4 * lane;     // rows of the current tile
    const float* ybase = p.y_true + (size_t)min(gwarp, ntiles - 1) * 32 * n + 4 * lane;
    if (gwarp < ntiles) {
#pragma unroll
        for (int i = 0; i < KEY_BATCH; ++i) issue_row(s4[i], y4[i], sbase, ybase, (gwarp << 5) + i, (unsigned)(i * n));
    }
    if ((int)threadIdx.x < (TPQ_K - 1) * 8) {
        const int r = (int)threadIdx.x >> 3, l = (int)threadIdx.x & 7;
        if (r < p.kmax && l <= 4) {
            if (F64) term64[threadIdx.x] = (p.exp_gains ? (double)((1 << l) - 1) : (double)l) / p.disc64[r];
            else term32[threadIdx.x] = (float)((1 << l) - 1) / p.disc32[r];
        }
    }
    __syncthreads();

    for (int t = gwarp; t < ntiles; t += nwarps) {
        const int q0 = t << 5;
        const bool more = t + nwarps < ntiles;
        const int qnext = more ? ((t + nwarps) << 5) : p.B;     // this warp's next tile (or nothing)
        const float* snext = p.y_pred + (size_t)(more ? qnext : q0) * n + 4 * lane;
        const float* ynext = p.y_true + (size_t)(more ? qnext : q0) * n + 4 * lane;
        // ---------------- phase 1: sorted quad keys -> shared memory, per-row histogram; tile-wide validity -------
        unsigned myhist = 0;
        unsigned err = 0;
        float ymin = 0.f, ymax = 0.f, amax = 0.f;
        auto do_row = [&](int r, const float4& sc, const float4& yc) {
            const float ss[4] = {sc.x, sc.y, sc.z, sc.w};
            const float ys[4] = {yc.x, yc.y, yc.z, yc.w};
            float k[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float tt = ys[c] + MAGIC;
                // the round trip returns the same bits iff the label is an integer (|y| < 2^22; -0 fails too)
                err |= __float_as_uint(tt - MAGIC) ^ __float_as_uint(ys[c]);
                k[c] = key_of(ss[c], tt);
            }
            ymin = fminf(fminf(ys[0], ys[1]), ymin); ymin = fminf(fminf(ys[2], ys[3]), ymin);
            ymax = fmaxf(fmaxf(ys[0], ys[1]), ymax); ymax = fmaxf(fmaxf(ys[2], ys[3]), ymax);
            amax = max_nan(max_nan(fabsf(ss[0]), fabsf(ss[1])), amax);
            amax = max_nan(max_nan(fabsf(ss[2]), fabsf(ss[3])), amax);
            const float code = fmaf(ys[3], 125.f, fmaf(ys[2], 25.f, fmaf(ys[1], 5.f, ys[0] + MAGIC)));
            const unsigned hist = lut[__float_as_uint(code) & 1023u];
            // the quad's three largest keys (sort of 4: 5 compare-exchanges, the smallest output unused)
            const float h01 = fmaxf(k[0], k[1]), l01 = fminf(k[0], k[1]);
            const float h23 = fmaxf(k[2], k[3]), l23 = fminf(k[2], k[3]);
            const float ka = fmaxf(h01, h23), m1 = fminf(h01, h23);
            const float m2 = fmaxf(l01, l23);
            const float kb = fmaxf(m1, m2), kc = fminf(m1, m2);
            tile.a[r * KEY_AB_STRIDE + lane] = ka;
            tile.b[r * KEY_AB_STRIDE + lane] = kb;
            tile.c[r * KEY_C_STRIDE + lane] = kc;
            const unsigned rowhist = __reduce_add_sync(LTR_FULL_MASK, hist);
            myhist = (lane == r) ? rowhist : myhist;
        };
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += KEY_BATCH) {
            // rows entering L2 now: KEY_AHEAD + KEY_BATCH rows ahead of the row in hand, in this tile or the warp's next
            const int ahead = rb + KEY_AHEAD + KEY_BATCH;
            prefetch_rows(ahead < 32 ? q0 + ahead : qnext + ahead - 32);
            // the slot a row leaves is reissued with the row KEY_BATCH further on, or a first row of the next tile
            const bool last = rb + KEY_BATCH >= 32;
            const float* sb = last ? snext : sbase;
            const float* yb = last ? ynext : ybase;
            const int qb = last ? qnext : q0 + rb + KEY_BATCH;
            const unsigned ob = last ? 0u : (unsigned)((rb + KEY_BATCH) * n);
#pragma unroll
            for (int i = 0; i < KEY_BATCH; ++i) {
                do_row(rb + i, s4[i], y4[i]);
                issue_row(s4[i], y4[i], sb, yb, qb + i, ob + (unsigned)(i * n));
            }
        }
        sbase = snext;
        ybase = ynext;
        // a label that is not an integer 0..4 or a NaN / infinite score anywhere in the tile: its 32 queries take the full sort
        const bool tile_ok = (err == 0u) && (ymin >= 0.f) && (ymax <= 4.0f) && (amax < finf);
        const int mybad = __any_sync(LTR_FULL_MASK, !tile_ok) ? 1 : 0;

        // ---------------- phase 2: one query per lane --------------------------------------------------
        const int q = q0 + lane;
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 m4 = *reinterpret_cast<const float4*>(&tile.a[lane * KEY_AB_STRIDE + 4 * j]);
            v[4 * j] = m4.x; v[4 * j + 1] = m4.y; v[4 * j + 2] = m4.z; v[4 * j + 3] = m4.w;
        }
#define TPQ_CEP(a, b)                         \
    {                                         \
        const float hi_ = fmaxf(v[a], v[b]);  \
        v[b] = fminf(v[a], v[b]);             \
        v[a] = hi_;                           \
    }
#include "ndcg_top11_32.inc"
#undef TPQ_CEP
        const float L = v[TPQ_K - 1];
        float top[TPQ_K];
#pragma unroll
        for (int j = 0; j < TPQ_K; ++j) top[j] = v[j];
        // quads whose second key lies above L: bit 31 - j
        unsigned extras = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(&tile.b[lane * KEY_AB_STRIDE + 4 * j]);
            extras = __funnelshift_l(__float_as_uint(L - b4.x), extras, 1);
            extras = __funnelshift_l(__float_as_uint(L - b4.y), extras, 1);
            extras = __funnelshift_l(__float_as_uint(L - b4.z), extras, 1);
            extras = __funnelshift_l(__float_as_uint(L - b4.w), extras, 1);
        }
        while (extras) {
            const int j = __clz((int)extras);
            extras &= ~(0x80000000u >> j);
            key_insert<TPQ_K>(top, tile.b[lane * KEY_AB_STRIDE + j]);
            const float kc = tile.c[lane * KEY_C_STRIDE + j];
            if (kc > L) {
                key_insert<TPQ_K>(top, kc);
                // three docs of one quad above L: its fourth key is rebuilt from the row (L2)
                const int qq = min(q, p.B - 1);
                const float4 sx = *reinterpret_cast<const float4*>(p.y_pred + (size_t)qq * n + 4 * j);
                const float4 yx = *reinterpret_cast<const float4*>(p.y_true + (size_t)qq * n + 4 * j);
                const float kd = fminf(fminf(key_of(sx.x, yx.x + MAGIC), key_of(sx.y, yx.y + MAGIC)),
                                       fminf(key_of(sx.z, yx.z + MAGIC), key_of(sx.w, yx.w + MAGIC)));
                if (kd > L) key_insert<TPQ_K>(top, kd);
            }
        }
        if (p.partial) {
            // a chunk of a long row: its 11 largest keys, label histogram and flag go to the merge (ndcg_chunk_merge_kernel)
            if (q < p.B) {
#pragma unroll
                for (int j = 0; j < TPQ_K; ++j) p.partial[(size_t)j * p.B + q] = top[j];
                p.partial[(size_t)TPQ_K * p.B + q] = __uint_as_float(myhist);
                p.partial[(size_t)(TPQ_K + 1) * p.B + q] = __uint_as_float((unsigned)mybad);
            }
            __syncwarp();                                        // the tile is free for the next phase 1
            continue;
        }
        // keys that agree in their upper 29 bits among the ranks the sweep reads: exact compares on the row itself
        int bad = mybad;
        int tied = 0;                                    // an exact tie seen by the exact path below
        bool near = false;
#pragma unroll
        for (int j = 0; j + 1 < TPQ_K; ++j) near = near || (j < p.kmax && j + 1 < n && key_near(top[j], top[j + 1]));
        // (the warp stages such a row in its tile, which no lane needs any more, and the query's lane scans it)
        __syncwarp();
        unsigned redo = __ballot_sync(LTR_FULL_MASK, near && !bad && q < p.B);
        if (__popc(redo) > 4) {                  // tie-heavy scores: no point in looking, the full sort takes them all
            bad |= (int)((redo >> lane) & 1u);
            redo = 0;
        }
        while (redo) {
            const int f = __ffs((int)redo) - 1;
            redo &= redo - 1;
            if (lane_in) {
                const size_t off = (size_t)(q0 + f) * n + 4 * lane;
                *reinterpret_cast<float4*>(&tile.a[4 * lane]) = *reinterpret_cast<const float4*>(p.y_pred + off);
                *reinterpret_cast<float4*>(&tile.b[4 * lane]) = *reinterpret_cast<const float4*>(p.y_true + off);
            }
            __syncwarp();
            if (lane == f) {
                const float ninf = __int_as_float(0xff800000);
                float xs[TPQ_K];
                int xl[TPQ_K];
#pragma unroll
                for (int j = 0; j < TPQ_K; ++j) xs[j] = ninf, xl[j] = 0;
#pragma unroll 1
                for (int d = 0; d < n; d += 4) {
                    const float4 sx = *reinterpret_cast<const float4*>(&tile.a[d]);
                    const float4 yx = *reinterpret_cast<const float4*>(&tile.b[d]);
                    const float sv[4] = {sx.x, sx.y, sx.z, sx.w}, yv[4] = {yx.x, yx.y, yx.z, yx.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float x = sv[c] + 0.0f;
                        if (x > xs[TPQ_K - 1]) tpq_insert<TPQ_K>(xs, xl, x, (int)yv[c]);
                    }
                }
#pragma unroll
                for (int j = 0; j < TPQ_K; ++j) top[j] = __uint_as_float((unsigned)xl[j]);     // only the label bits are read below
#pragma unroll
                for (int j = 0; j + 1 < TPQ_K; ++j) tied |= (j < p.kmax && j + 1 < n && xs[j] == xs[j + 1]) ? 1 : 0;
                bad |= tied;
            }
            __syncwarp();
        }
        // a marked query (exact tie among the ranks the sweep reads, label outside 0..4, NaN / infinite score) is sorted in
        // full right here, by the warp, in the tile it no longer needs: bitonic network + std::sort replay (sort_common.cuh)
        {
            __syncwarp();
            unsigned todo = __ballot_sync(LTR_FULL_MASK, bad && q < p.B);
            const unsigned ties = __ballot_sync(LTR_FULL_MASK, tied);       // these go straight to the std::sort replay
            while (todo) {
                const int f = __ffs((int)todo) - 1;
                todo &= todo - 1;
                // (a tile that passed the label check has its rows' histograms: the ideal order needs no second sort)
                key_full_sort<F64>(p, tile.a, q0 + f, (ties >> f) & 1u, mybad == 0, __shfl_sync(LTR_FULL_MASK, myhist, f));
            }
        }
        if (q < p.B) {
            p.flags[1 + q] = bad;
            if (!bad) {
                // ideal order from the histogram: level L fills the ranks below #docs with label >= L.  One 3-bit field
                // per rank: the sum over the levels of (rank < #docs with label >= L)
                const int c4 = myhist >> 24, c3 = c4 + ((myhist >> 16) & 255), c2 = c3 + ((myhist >> 8) & 255),
                          c1 = c2 + (myhist & 255);
                const unsigned ONES = 0x09249249u;                       // bit 3 r, r < 10
                const unsigned ideal_lab = (((1u << (3 * min(c4, 10))) - 1u) & ONES) + (((1u << (3 * min(c3, 10))) - 1u) & ONES) +
                                           (((1u << (3 * min(c2, 10))) - 1u) & ONES) + (((1u << (3 * min(c1, 10))) - 1u) & ONES);
                unsigned pred_lab = 0;
#pragma unroll
                for (int r = 0; r < TPQ_K - 1; ++r) pred_lab |= (__float_as_uint(top[r]) & 7u) << (3 * r);
                if (!F64) {
                    float got = 0.f, ideal = 0.f;
                    float* out = p.out32 + (size_t)q * p.nk;
#pragma unroll 1
                    for (int r = 0; r < p.kmax; ++r) {
                        got += term32[8 * r + ((pred_lab >> (3 * r)) & 7u)];
                        ideal += term32[8 * r + ((ideal_lab >> (3 * r)) & 7u)];
                        if ((kends >> r) & 1u) {
                            const float val = got / ideal;
                            for (int j = 0; j < p.nk; ++j)
                                if (p.ks[j] == r + 1) out[j] = val;
                        }
                    }
                } else {
                    double got = 0.0, ideal = 0.0;
                    double* out = p.out64 + (size_t)q * p.nk;
#pragma unroll 1
                    for (int r = 0; r < p.kmax; ++r) {
                        got += term64[8 * r + ((pred_lab >> (3 * r)) & 7u)];
                        ideal += term64[8 * r + ((ideal_lab >> (3 * r)) & 7u)];
                        if ((kends >> r) & 1u) {
                            const double val = p.raw_dcg ? got : ((ideal == 0.0) ? p.empty_value : got / ideal);
                            for (int j = 0; j < p.nk; ++j)
                                if (p.ks[j] == r + 1) out[j] = val;
                        }
                    }
                }
            }
        }
        __syncwarp();                                            // the tile is free for the next phase 1
    }
}

// Rows of several whole chunks: ndcg_key_kernel has left, per chunk, the 11 largest keys (label in the low 3 bits), the label
// histogram and the flag in p.partial ([13][B * chunks], chunk c of query q at column q * chunks + c).  A group of G = 2..32 lanes
// (the chunk count rounded up to a power of two) owns a query: lane c loads chunk c's list (coalesced: neighbouring lanes read
// neighbouring columns), then log2 G butterfly rounds merge the lists -- a lane inserts its partner's 11 keys, 2 FMNMX per slot and
// key, so every lane of the group ends with the query's 11 largest keys -- and add the histograms.  The group's first lane
// finishes like the key kernel; keys that agree in their upper 29 bits among the ranks the sweep reads (an exact tie, or scores
// 1e-7 apart) are left to ndcg_kernel, which replays the reference's sort on the row (flag 2), as are rows with odd labels or
// non-finite scores (flag 1).
__global__ void __launch_bounds__(128) ndcg_chunk_merge_kernel(const NdcgParams p) {
    __shared__ float term32[(TPQ_K - 1) * 8];
    const int chunks = p.n / TPQ_N;
    const size_t rows = (size_t)p.B * chunks;
    int G = 2;
    while (G < chunks && G < 32) G <<= 1;
    unsigned kends = 0;                                    // bit r: some k of the sweep equals r + 1
    for (int j = 0; j < p.nk; ++j) kends |= 1u << (p.ks[j] - 1);
    pdl_prologue();
    if ((int)threadIdx.x < (TPQ_K - 1) * 8) {
        const int r = (int)threadIdx.x >> 3, l = (int)threadIdx.x & 7;
        if (r < p.kmax && l <= 4) term32[threadIdx.x] = (float)((1 << l) - 1) / p.disc32[r];
    }
    __syncthreads();
    const int per_cta = 128 / G;                           // queries per CTA
    const int sub = (int)threadIdx.x % G;                  // lane of the group (groups do not straddle warps: G <= 32)
    const int q = (int)blockIdx.x * per_cta + (int)threadIdx.x / G;
    const bool live = q < p.B;                             // (whole groups are live or not: the shuffles below stay convergent)
    float top[TPQ_K];
#pragma unroll
    for (int j = 0; j < TPQ_K; ++j) top[j] = __uint_as_float(0xff7ffff8u);      // below every finite score, label 0
    int c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    unsigned bad = 0;
    if (live) {
#pragma unroll 1
        for (int c = sub; c < chunks; c += G) {            // (more than 32 chunks: a lane folds several)
            const size_t col = (size_t)q * chunks + c;
            float mine[TPQ_K];
#pragma unroll
            for (int j = 0; j < TPQ_K; ++j) mine[j] = __ldcg(p.partial + (size_t)j * rows + col);
            const unsigned h = __float_as_uint(__ldcg(p.partial + (size_t)TPQ_K * rows + col));
            bad |= __float_as_uint(__ldcg(p.partial + (size_t)(TPQ_K + 1) * rows + col));
#pragma unroll
            for (int j = 0; j < TPQ_K; ++j) key_insert<TPQ_K>(top, mine[j]);
            c4 += (int)(h >> 24), c3 += (int)((h >> 16) & 255u), c2 += (int)((h >> 8) & 255u), c1 += (int)(h & 255u);
        }
    }
    for (int o = 1; o < G; o <<= 1) {
        float theirs[TPQ_K];
#pragma unroll
        for (int j = 0; j < TPQ_K; ++j) theirs[j] = __shfl_xor_sync(LTR_FULL_MASK, top[j], o);
#pragma unroll
        for (int j = 0; j < TPQ_K; ++j) key_insert<TPQ_K>(top, theirs[j]);
        c1 += __shfl_xor_sync(LTR_FULL_MASK, c1, o);
        c2 += __shfl_xor_sync(LTR_FULL_MASK, c2, o);
        c3 += __shfl_xor_sync(LTR_FULL_MASK, c3, o);
        c4 += __shfl_xor_sync(LTR_FULL_MASK, c4, o);
        bad |= __shfl_xor_sync(LTR_FULL_MASK, bad, o);
    }
    if (!live || sub != 0) return;
    bool near = false;
#pragma unroll
    for (int j = 0; j + 1 < TPQ_K; ++j) near = near || (j < p.kmax && key_near(top[j], top[j + 1]));
    p.flags[1 + q] = bad ? 1 : (near ? 2 : 0);
    if (bad || near) return;
    c3 += c4, c2 += c3, c1 += c2;                              // #docs with label >= L
    unsigned pred_lab = 0;                                     // 3 bits per rank
#pragma unroll
    for (int r = 0; r < TPQ_K - 1; ++r) pred_lab |= (__float_as_uint(top[r]) & 7u) << (3 * r);
    float got = 0.f, ideal = 0.f;
    float* out = p.out32 + (size_t)q * p.nk;
#pragma unroll 1
    for (int r = 0; r < p.kmax; ++r) {
        const int li = (c4 > r) + (c3 > r) + (c2 > r) + (c1 > r);
        got += term32[8 * r + ((pred_lab >> (3 * r)) & 7u)];
        ideal += term32[8 * r + li];
        if ((kends >> r) & 1u) {
            const float val = got / ideal;
            for (int j = 0; j < p.nk; ++j)
                if (p.ks[j] == r + 1) out[j] = val;
        }
    }
}

}  // namespace ltr
