// K2: fused forward+backward LambdaLoss framework (replaces losses/lambdaL.py:7-127 of the reference):
// lambdaMask + lambdaLoss with the seven weighing schemes, NDCG@k truncation, sum / mean reduction and
// binary / natural logarithm.  One CTA per slate:
//   1. scores (padding -> -inf) are sorted in shared memory with torch.sort's tie order (sort_common.cuh),
//      labels follow the permutation; a second, values-only sort of the labels gives the ideal DCG@k;
//   2. each thread owns rows i = tid, tid + 128, ... of the k x k block of rank positions and walks the
//      columns j; a visited pair (i, j) with label_i != label_j belongs to the mask exactly once, as
//      (hi, lo) = the doc with the larger label first, so ONE evaluation per visit serves both the
//      loss (counted by the thread that owns `hi`) and d loss / d s_i;
//   3. per pair: e = 2^(-sigma log2e (s_hi - s_lo)), sig = 1/(1+e), l = w * max(log2 eps, -log2(1+e)),
//      clamped again at log2 eps (lambdaL.py:51), d l / d delta = w sigma (1 - sig) where unclamped:
//      3 MUFU (ex2, rcp, lg2) + ~15 FP32; the B x n x n tensors of the reference never exist.
// `ndcgLoss1_scheme` has no label condition (lambdaL.py:26): every ordered pair of real docs, the diagonal
// included, is in the mask with weight G_i / D_i.
#include <stdlib.h>

#include "lambda_common.cuh"
#include "ltr_host.h"
#include "sort_common.cuh"

namespace ltr {

constexpr int LAMBDA_THREADS = 128;
constexpr int LAMBDA_MAX_N = 2048;

struct LambdaParams {
    const float* y_pred;
    const float* y_true;
    int B, n, len, k;           // k < 0: no truncation; k = 0: nothing inside the truncation (lambdaL.py:29-30)
    float eps, pad, sigma, mu;
    int natural_log;            // 0: log2, 1: ln
    float* slate_sum;           // [B] sum of the masked pair log-probabilities of the slate
    float* slate_cnt;           // [B] number of masked pairs
    float* grad;                // [B, n] d(-sum)/d y_pred, unscaled; null = forward only
    unsigned* next;             // next-slate counter in the caller's workspace (left at 0), or null: static assignment
    // longest-slate-first schedule (lambda_order_kernel), or null: slates in index order
    const unsigned* order;      // [B] slate indices by decreasing pair work
    unsigned* claimed;          // [B] one-wave form: claim flag of order[i]
    unsigned* arrive;           // [LAMBDA_ARRIVE] one-wave form: CTAs that have started on SM (smid % nsm)
    unsigned* cursor;           // one-wave form: steal cursor (from the short end of `order`)
    int one_wave, nsm;          // every CTA takes exactly one slate (grid == B, all resident) / SMs of the device
    // the last CTA to leave reduces the slates' sums (no second launch), or null: lambda_finish_kernel does
    unsigned* finished;         // exited-CTA counter in the workspace (left at 0)
    float* loss;                // [1] -sum, or -sum / count for reduction "mean"
    float* count_out;           // [1] number of masked pairs, or null
    int mean;
    int force_track;            // debugging: never take the untracked tiles
};
constexpr int LAMBDA_ORDER_MAX_B = 4096;        // batches the schedule is built for (its arrays live in the workspace)
constexpr int LAMBDA_ARRIVE = 1024;
constexpr int LAMBDA_KEYS = 512;                // work classes: block lengths 0..256 one by one, longer ones by eights

// ---- symmetric pair schedule for top-k blocks of up to 256 docs --------------------------------------
// Same block-circulant ownership as the approxNDCG fast path: lane a of every warp owns R consecutive
// rank positions (score, label, gain, 1/D in registers); at step D it meets the R docs of block (a + D) mod T,
// which travel from lane to lane by warp shuffle together with their gradient accumulators.  Every
// UNORDERED pair is visited once (twice at weight 1/2 on the antipodal step of an even T): the doc with
// the larger label is the "hi" side, one evaluation gives the loss term and both gradient contributions.
// Steps are split over the 4 warps; partial gradients meet in shared memory.
struct LDoc {
    float s, v, y, g, d;   // score (FACT: u = 2^(t - c)), v = 2^-(t - c) (FACT only), label (NaN: not a doc of the
                           // block), gain, 1 / log2(1 + rank);  t = sigma log2(e) s
};

// The pair tiles run WITHOUT the two eps clamps of lambdaL.py:51 and only track the smallest log2-sigmoid (mn) and the
// smallest weighted one (mn2) they produced: a slate none of whose pairs went below log2(eps) -- every slate of real
// data with the default eps = 1e-10 -- has had no active clamp, so its sums are the reference's; a slate that did is done
// again by the exact row-per-thread path (the kernel checks mn / mn2 once per slate).  Costs two FMNMX per pair instead
// of two clamps, two compares and a select.
// log2-probability of one pair from e = exp(-sigma (s_first - s_second)) = v_first u_second and its derivative
// (lambdaL.py:51-55 without the final log-base factor, which the caller applies once per slate)
__device__ __forceinline__ void pair_from_e(float e, float w, float& lw, float& dl, float& mn, float& mn2) {
    const float q = 1.0f + e;
    const float sig = rcp_approx(q);
    const float l2 = -lg2_approx(q);                              // log2 sigmoid
    lw = w * l2;                                                  // log2 of (.)^w
    mn = fminf(mn, l2);
    mn2 = fminf(mn2, lw);
    dl = w * (e * sig);                                           // w sigma (1 - sig); 1 - sig = e sig
}

// One unordered pair of the block in the factorised form (u = 2^(t - c) in .s, v = 2^-(t - c) in .v; docs outside the block
// carry u = v = 0 and y = NaN: e = 0, so whatever their weight, they add 0 to the loss and to both gradients).
// Equal labels: the gain-based weights and the label-difference weights are 0 by themselves; only the constant weight of
// `None` needs the label comparison.  The number of masked pairs is not counted here (closed form, see the kernel).
// TRACK = false (the kernel has bounded the slate's log2-sigmoids and pair weights from its score range and label range: no
// clamp can become active, nothing to track) and a gain-difference weight: the SIGNED weight ws = h (g_a - g_b)(mu dt + |dd|)
// carries which doc is the better-labelled one (gains grow with the label), so the loss term is |ws| log2 sig and the
// gradient pair is -+ ws (1 - sig), with no label in the tile at all: 16 instructions per pair instead of 21.
template <int SCHEME>
constexpr bool lambda_gain_scheme() { return SCHEME == SCHEME_NDCG2 || SCHEME == SCHEME_LAMBDARANK || SCHEME == SCHEME_NDCG2PP; }

template <int SCHEME, bool TRACK>
__device__ __forceinline__ void lambda_pair(const LDoc& a, const LDoc& b, float h, float dt, float mu, float& loss, float& mn,
                                            float& mn2, float& ga, float& gb) {
    if (!TRACK && lambda_gain_scheme<SCHEME>()) {
        const float dgs = a.g * h - b.g * h;                   // (h = 1 or 1/2: exact; the products are shared by the tile's pairs)
        float t;
        if (SCHEME == SCHEME_NDCG2) t = dt;
        else if (SCHEME == SCHEME_LAMBDARANK) t = fabsf(a.d - b.d);
        else t = fmaf(mu, dt, fabsf(a.d - b.d));
        const float ws = dgs * t;
        const float e = dgs > 0.f ? a.v * b.s : b.v * a.s;     // equal gains: ws = 0, either e will do
        const float q = 1.0f + e;
        const float sig = rcp_approx(q);
        loss = fmaf(fabsf(ws), -lg2_approx(q), loss);
        const float dl = ws * (e * sig);
        ga -= dl;
        gb += dl;
        return;
    }
    if (SCHEME == SCHEME_NDCG1) {
        if (!(a.y == a.y) || !(b.y == b.y)) return;
        // no label condition: both ordered pairs are in the mask, weight G_i / D_i of the FIRST doc (lambdaL.py:96-97)
        float l1, d1, l2, d2;
        pair_from_e(a.v * b.s, a.g * a.d, l1, d1, mn, mn2);
        pair_from_e(b.v * a.s, b.g * b.d, l2, d2, mn, mn2);
        loss = fmaf(h, l1 + l2, loss);
        const float dd = h * (d2 - d1);
        ga += dd;
        gb -= dd;
        return;
    }
    const bool a_hi = a.y > b.y;
    float w;
    if (SCHEME == SCHEME_NONE) w = (a_hi || (a.y < b.y)) ? 1.0f : 0.f;
    else if (SCHEME == SCHEME_GTDIFF) w = fabsf(fmaxf(a.y, 0.f) - fmaxf(b.y, 0.f));
    else if (SCHEME == SCHEME_GTDIFF_POW) {
        const float u = fmaxf(a.y, 0.f), v = fmaxf(b.y, 0.f);
        w = fabsf(u * u - v * v);
    } else {
        // dt = |1/D_t - 1/D_{t+1}| at the pair's rank gap t: loaded by the caller once per tile diagonal (the gap of (row r,
        // column c) of a tile depends on r - c only), not looked up per pair
        const float dg = fabsf(a.g - b.g);
        if (SCHEME == SCHEME_NDCG2) w = dt * dg;
        else if (SCHEME == SCHEME_LAMBDARANK) w = fabsf(a.d - b.d) * dg;
        else w = dg * fmaf(mu, dt, fabsf(a.d - b.d));
    }
    // the better-labelled doc is the first one of the pair
    const float e = a_hi ? a.v * b.s : b.v * a.s;
    float lw, dl;
    pair_from_e(e, w, lw, dl, mn, mn2);
    loss = fmaf(h, lw, loss);
    dl *= a_hi ? -h : h;                      // d(-sum l)/d s_hi = -dl, d/d s_lo = +dl
    ga += dl;
    gb -= dl;
}

// dtab values of a tile whose rows start at rank `r0` and whose columns start at rank `c0` (different blocks: |r0 - c0| >= R):
// dv[r - c + R - 1] = dtab[|r0 + r - (c0 + c)|].  Only the schemes with a rank-gap weight read them.
template <int SCHEME, int R>
__device__ __forceinline__ void tile_dtab(const float* dtab, int r0, int c0, float (&dv)[2 * R - 1]) {
    if (SCHEME == SCHEME_NDCG2 || SCHEME == SCHEME_NDCG2PP) {
        const int bd = r0 - c0;
        const int s0 = bd < 0 ? -bd : bd;
#pragma unroll
        for (int k = 0; k < 2 * R - 1; ++k) {
            const int off = bd < 0 ? (R - 1 - k) : (k - (R - 1));
            dv[k] = dtab[min(s0 + off, 255)];                  // (lanes beyond the block read anything: their pairs are masked)
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2 * R - 1; ++k) dv[k] = 0.f;
    }
}

// K, Y, G, invD: the slate in rank order (Y = -inf beyond the block); gpart: [4][2][256] floats of shared memory.
// On exit gpart holds the per-warp row / column partial gradients of positions 0 .. lim-1 (after a barrier).
// K holds u = 2^(t - c) and V holds 2^-(t - c) of the block's docs.  mn / mn2: see pair_from_e.
template <int SCHEME, int R, bool TRACK>
__device__ __forceinline__ void lambda_sym_slate(const float* K, const float* V, const float* Y, const float* G,
                                                 const float* invD, const float* dtab, int lim, int b, float mu,
                                                 float* gpart, float& loss, float& mn, float& mn2) {
    constexpr bool FACT = true;
    const float ninf = __int_as_float(0x7fc00000);           // (a quiet NaN marks "not a doc of the block")
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = (lim + R - 1) / R;
    const int S = T >> 1;
    const int role = (warp + b) & 3;
    const int d_lo = 1 + ((role * S) >> 2);
    const int d_hi = 1 + (((role + 1) * S) >> 2);
    const bool active = lane < T;
    const int src = active ? ((lane + 1 == T) ? 0 : lane + 1) : lane;
    int cb = lane + d_lo;
    if (cb >= T) cb -= T;
    if (!active) cb = 0;
    LDoc row[R], col[R];
    float g_r[R], g_c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = lane * R + r, j = cb * R + r;
        const bool vi = active && i < lim, vj = active && j < lim;
        row[r].s = vi ? K[i] : 0.f; row[r].y = vi ? Y[i] : ninf; row[r].g = vi ? G[i] : 0.f; row[r].d = vi ? invD[i] : 0.f;
        col[r].s = vj ? K[j] : 0.f; col[r].y = vj ? Y[j] : ninf; col[r].g = vj ? G[j] : 0.f; col[r].d = vj ? invD[j] : 0.f;
        row[r].v = (FACT && vi) ? V[i] : 0.f;
        col[r].v = (FACT && vj) ? V[j] : 0.f;
        g_r[r] = 0.f;
        g_c[r] = 0.f;
    }
    if (role == 0) {                          // pairs inside the lane's own block (+ the diagonal of ndcgLoss1)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (SCHEME == SCHEME_NDCG1 && row[r].y == row[r].y) {
                float l, dl;
                pair_from_e(1.0f, row[r].g * row[r].d, l, dl, mn, mn2);    // the diagonal: delta = 0
                loss += l;
            }
#pragma unroll
            for (int c = r + 1; c < R; ++c)
                lambda_pair<SCHEME, TRACK>(row[r], row[c], 1.0f, dtab[c - r], mu, loss, mn, mn2, g_r[r], g_r[c]);
        }
    }
#pragma unroll 1
    for (int D = d_lo; D < d_hi; ++D) {
        const float h = (2 * D == T) ? 0.5f : 1.0f;
        int cblk = lane + D;                   // block visiting this lane at step D
        if (cblk >= T) cblk -= T;
        const int j0 = cblk * R;
        float dv[2 * R - 1];
        tile_dtab<SCHEME, R>(dtab, lane * R, j0, dv);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = 0; c < R; ++c)
                lambda_pair<SCHEME, TRACK>(row[r], col[c], h, dv[r - c + R - 1], mu, loss, mn, mn2, g_r[r], g_c[c]);
        }
        if (D + 1 < d_hi) {
#pragma unroll
            for (int c = 0; c < R; ++c) {
                col[c].s = __shfl_sync(LTR_FULL_MASK, col[c].s, src);
                if (FACT) col[c].v = __shfl_sync(LTR_FULL_MASK, col[c].v, src);
                if (TRACK || !lambda_gain_scheme<SCHEME>()) col[c].y = __shfl_sync(LTR_FULL_MASK, col[c].y, src);
                col[c].g = __shfl_sync(LTR_FULL_MASK, col[c].g, src);
                col[c].d = __shfl_sync(LTR_FULL_MASK, col[c].d, src);
                g_c[c] = __shfl_sync(LTR_FULL_MASK, g_c[c], src);
            }
        }
    }
    if (active) {
        int cblk = lane + d_hi - 1;           // the block whose accumulators ended up here
        if (cblk >= T) cblk -= T;
        if (d_hi == d_lo) cblk = lane;        // nothing accumulated: zeros to the own slots
#pragma unroll
        for (int r = 0; r < R; ++r) {
            gpart[(warp * 2 + 0) * 256 + lane * R + r] = g_r[r];
            gpart[(warp * 2 + 1) * 256 + cblk * R + r] = g_c[r];
        }
    }
    __syncthreads();
}

// Blocks of 65..256 docs with 4 docs per lane (80-odd registers, one small pair tile that stays in the instruction
// cache): the rank positions are split into group A = [0, 128) and group B = [128, lim).  Three circulant phases run
// the SAME tile code:  A x A and B x B are symmetric (steps 1..T/2 of a ring of T = ceil(group/4) lanes + the
// in-block triangle);  A x B is a full ring of 32 lanes, rows from A, columns from B (blocks beyond B are non-docs).
// The steps of every phase are split over the 4 warps; a lane's row accumulators stay in registers across phases.
template <int SCHEME, bool TRACK>
__device__ __forceinline__ void lambda_sym_slate4(const float* K, const float* V, const float* Y, const float* G,
                                                  const float* invD, const float* dtab, int lim, int b, float mu,
                                                  float* gpart, float& loss, float& mn, float& mn2) {
    constexpr int R = 4;
    constexpr bool FACT = true;
    const float qnan = __int_as_float(0x7fc00000);            // (a quiet NaN marks "not a doc of the block")
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int role = (warp + b) & 3;
    const int nA = lim < 128 ? lim : 128;
    const int TA = (nA + R - 1) / R, TB = (lim - nA + R - 1) / R;
    auto load_doc = [&](LDoc& d, int i) {
        const bool ok = i < lim;
        d.s = ok ? K[i] : 0.f;
        d.v = (FACT && ok) ? V[i] : 0.f;
        d.y = ok ? Y[i] : qnan;
        d.g = ok ? G[i] : 0.f;
        d.d = ok ? invD[i] : 0.f;
    };
    // this warp's column partials are accumulated over the phases: start from zero
    float* colpart = gpart + (warp * 2 + 1) * 256;
    float* rowpart = gpart + (warp * 2 + 0) * 256;
#pragma unroll
    for (int k = 0; k < 8; ++k) colpart[lane + 32 * k] = 0.f;
    __syncwarp();
    LDoc row[R], col[R];
    float g_r[R], g_c[R];
    // phases in the order A x A, A x B, B x B: the row group changes once
#pragma unroll 1
    for (int phase = 0; phase < 3; ++phase) {
        if (phase > 0 && TB == 0) break;
        const bool cross = phase == 1;
        // ring size.  A x B: the smallest power of two >= TB; the 32 lanes form 32 / T independent rings, each carrying
        // its own copy of group B's blocks past its lanes' rows, so no step is spent on blocks that do not exist
        int T = phase == 0 ? TA : TB;
        if (cross) {
            T = 1;
            while (T < TB) T <<= 1;
        }
        const int rbase = phase == 2 ? 128 : 0;                 // first rank position of the row group
        const int cbase = phase == 0 ? 0 : 128;                 //                       of the column group
        const int TR = phase == 2 ? TB : TA;                    // lanes that own a row block
        const int S = T >> 1;
        // symmetric: steps 1 .. S; cross: steps 0 .. T - 1
        const int d_lo = cross ? (role * T) >> 2 : 1 + ((role * S) >> 2);
        const int d_hi = cross ? ((role + 1) * T) >> 2 : 1 + (((role + 1) * S) >> 2);
        const bool active = cross || lane < T;
        const int ring0 = cross ? (lane & ~(T - 1)) : 0;        // first lane of this lane's ring
        const int pos = lane - ring0;                           // position inside the ring
        const int src = active ? ring0 + ((pos + 1 == T) ? 0 : pos + 1) : lane;
        int cb = pos + d_lo;
        if (cb >= T) cb -= T;
        if (!active) cb = 0;
        if (phase != 1) {                          // (A x B keeps the rows and row accumulators of A x A)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                load_doc(row[r], lane < TR ? rbase + lane * R + r : lim);
                g_r[r] = 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            load_doc(col[r], active ? cbase + cb * R + r : lim);
            g_c[r] = 0.f;
        }
        if (!cross && role == 0) {                // pairs inside the lane's own block (+ the diagonal of ndcgLoss1)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (SCHEME == SCHEME_NDCG1 && row[r].y == row[r].y) {
                    float l, dl;
                    pair_from_e(1.0f, row[r].g * row[r].d, l, dl, mn, mn2);    // the diagonal: delta = 0
                    loss += l;
                }
#pragma unroll
                for (int c = r + 1; c < R; ++c)
                    lambda_pair<SCHEME, TRACK>(row[r], row[c], 1.0f, dtab[c - r], mu, loss, mn, mn2, g_r[r], g_r[c]);
            }
        }
#pragma unroll 1
        for (int D = d_lo; D < d_hi; ++D) {
            const float h = (!cross && 2 * D == T) ? 0.5f : 1.0f;
            int cblk = pos + D;                    // block visiting this lane at step D
            if (cblk >= T) cblk -= T;
            const int j0 = cbase + cblk * R;
            float dv[2 * R - 1];
            tile_dtab<SCHEME, R>(dtab, rbase + lane * R, j0, dv);
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int c = 0; c < R; ++c)
                    lambda_pair<SCHEME, TRACK>(row[r], col[c], h, dv[r - c + R - 1], mu, loss, mn, mn2, g_r[r], g_c[c]);
            }
            if (D + 1 < d_hi) {
#pragma unroll
                for (int c = 0; c < R; ++c) {
                    col[c].s = __shfl_sync(LTR_FULL_MASK, col[c].s, src);
                    if (FACT) col[c].v = __shfl_sync(LTR_FULL_MASK, col[c].v, src);
                    if (TRACK || !lambda_gain_scheme<SCHEME>()) col[c].y = __shfl_sync(LTR_FULL_MASK, col[c].y, src);
                    col[c].g = __shfl_sync(LTR_FULL_MASK, col[c].g, src);
                    col[c].d = __shfl_sync(LTR_FULL_MASK, col[c].d, src);
                    g_c[c] = __shfl_sync(LTR_FULL_MASK, g_c[c], src);
                }
            }
        }
        // column partials of the block whose accumulators ended up in this lane (same warp: the phases are ordered,
        // and so are the rings of A x B, which all hold a copy of the same blocks)
        {
            int cblk = pos + d_hi - 1;
            if (cblk >= T) cblk -= T;
            const int nrings = cross ? 32 / T : 1;
            for (int ring = 0; ring < nrings; ++ring) {
                if (active && d_hi > d_lo && ring0 == ring * T && cblk < (cross ? TB : T)) {
#pragma unroll
                    for (int r = 0; r < R; ++r) colpart[cbase + cblk * R + r] += g_c[r];
                }
                __syncwarp();
            }
        }
        // row partials: group A after A x B (or after A x A when there is no group B), group B after B x B
        if ((phase == 1 || phase == 2 || TB == 0) && lane < TR) {
#pragma unroll
            for (int r = 0; r < R; ++r) rowpart[rbase + lane * R + r] = g_r[r];
        }
    }
    __syncthreads();
}

// The tracking tiles as a function of their own where the kernel also holds the untracked ones (the gain schemes): inlined next
// to them they cost the common path its register allocation (measured: 50 -> 56 us for the tracked path, 52 for the untracked).
template <int SCHEME>
__device__ __noinline__ void lambda_slate_tracked_call(const float* K, const float* V, const float* Y, const float* G, const float* invD,
                                                       const float* dtab, int lim, int b, float mu, float* gpart, float* acc) {
    float loss = acc[0], mn = acc[1], mn2 = acc[2];
    if (lim > 64) lambda_sym_slate4<SCHEME, true>(K, V, Y, G, invD, dtab, lim, b, mu, gpart, loss, mn, mn2);
    else lambda_sym_slate<SCHEME, 2, true>(K, V, Y, G, invD, dtab, lim, b, mu, gpart, loss, mn, mn2);
    acc[0] = loss, acc[1] = mn, acc[2] = mn2;
}
template <int SCHEME>
__device__ __forceinline__ void lambda_slate_tracked(const float* K, const float* V, const float* Y, const float* G, const float* invD,
                                                     const float* dtab, int lim, int b, float mu, float* gpart, float& loss,
                                                     float& mn, float& mn2) {
    if (lambda_gain_scheme<SCHEME>()) {
        float acc[3] = {loss, mn, mn2};
        lambda_slate_tracked_call<SCHEME>(K, V, Y, G, invD, dtab, lim, b, mu, gpart, acc);
        loss = acc[0], mn = acc[1], mn2 = acc[2];
    } else if (lim > 64) {
        lambda_sym_slate4<SCHEME, true>(K, V, Y, G, invD, dtab, lim, b, mu, gpart, loss, mn, mn2);
    } else {
        lambda_sym_slate<SCHEME, 2, true>(K, V, Y, G, invD, dtab, lim, b, mu, gpart, loss, mn, mn2);
    }
}

#ifndef LAMBDA_MIN_CTAS
#define LAMBDA_MIN_CTAS 7       // resident CTAs per SM the register allocation is tuned for (72 registers, 200 bytes of spills).
                                // Measured with the clamp-free pair tiles, us per call at 4 / 6 / 7 / 8 CTAs per SM:
                                //   1024 x 256: 58.0 / 58.9 / 53.0 / 56.1 (7 x 148 = 1036 slots: the batch is ONE wave)
                                //   4096 x 256: 225 / 170 / 180 / 188;   2048 x 128: 52.0 / 46.5 / 43.4 / 43.8;   1024 x 100: 28.8 / 28.3 / 24.5 / 24.3
                                // (4 was best for the round-1 code, whose tiles needed 128 registers)
                                // End of round 2 (untracked tiles, register sort, longest-first schedule), 1024 x 256: 44.6 / 46.5 / 44.1 / 46.5 at 5 / 6 / 7 / 8
#endif
template <int SCHEME>
__global__ void __launch_bounds__(LAMBDA_THREADS, LAMBDA_MIN_CTAS) lambda_kernel(const LambdaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[32];
    __shared__ int sh_m;
    __shared__ int sh_lvl[5];                               // [0]: a label that is not an integer 0..4; [L]: #docs with label L
    const int n = p.n, len = p.len;
    float* K = reinterpret_cast<float*>(smem_raw);          // scores in rank order
    int* I = reinterpret_cast<int*>(K + len);               // permutation
    float* Y = reinterpret_cast<float*>(I + len);           // sort scratch, then labels in rank order (-inf = padding)
    float* G = Y + len;                                     // ideal labels, then gains in rank order
    float* invD = G + len;                                  // 1 / log2(1 + r), r = 1..n
    float* dtab = invD + len;                               // |1/D_t - 1/D_{t+1}| by rank gap t (0 on the diagonal)
    float* gpart = dtab + len;                              // [4 warps][row / column][256] partial gradients (symmetric path)
    float* U = gpart + 8 * 256;                             // factorised pair exponent: 2^(t - c) of the block's docs
    float* V = U + 256;                                     //                           2^-(t - c)
    const float ninf = __int_as_float(0xff800000);
    const int tid = threadIdx.x;

    // (the tables depend on nothing the previous kernel on the stream wrote: built while it drains)
    for (int r = tid; r < n; r += LAMBDA_THREADS) invD[r] = 1.0f / log2f(2.0f + (float)r);
    __syncthreads();

    for (int t = tid; t < n; t += LAMBDA_THREADS) dtab[t] = (t == 0) ? 0.f : fabsf(invD[t - 1] - invD[t]);   // lambdaL.py:101-104
    pdl_prologue();
    const float sig_scale = p.sigma * kLog2e;
    const float log_eps2 = log2f(p.eps);
    const float out_scale = p.natural_log ? kLn2 : 1.0f;
    // d/d delta of log_b: w sigma (1-sig) / ln b -> binary: / ln 2, natural: 1
    const float grad_scale = p.natural_log ? p.sigma : p.sigma / kLn2;

    // slates cost up to 4x each other (the pair work grows with the square of their length): CTAs take the next slate
    // from a counter as they finish instead of a fixed share
    __shared__ int sh_next;
    int b = blockIdx.x;
    if (p.order) {
        // Longest first.  Several waves: the resident CTAs start on the longest slates and take the next-longest from the
        // counter as they finish, so the tail of the launch is made of the shortest ones.  One wave (every slate has its own
        // resident CTA, nothing can be handed out later): what finishes last is the SM with the most pair work, so the k-th
        // CTA to start on SM s takes entry k * nsm + (s or nsm - 1 - s, alternating with k) of the sorted list -- every SM gets
        // one slate of every length class, in snake order, and the per-SM totals agree to a percent instead of +-25 %.
        // Whatever the hardware's placement, claim flags keep it exact: an entry is taken once; a CTA whose entry does not
        // exist or is gone steals from the short end.
        if (tid == 0) {
            int slate = p.B;                                                       // (nothing left: leave)
            if (p.one_wave) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                const int s = (int)(smid % (unsigned)p.nsm);
                const unsigned k = atomicAdd(p.arrive + (s & (LAMBDA_ARRIVE - 1)), 1u);
                const long long pos = (long long)k * p.nsm + ((k & 1u) ? p.nsm - 1 - s : s);
                const unsigned mine = pos < p.B ? __ldcg(p.order + pos) : 0u;      // (in flight together with the claim)
                if (pos < p.B && atomicCAS(p.claimed + pos, 0u, 1u) == 0u) slate = (int)mine;
                while (slate == p.B) {
                    const unsigned q = atomicAdd(p.cursor, 1u);
                    if (q >= (unsigned)p.B) break;
                    const int at = p.B - 1 - (int)q;
                    if (atomicCAS(p.claimed + at, 0u, 1u) == 0u) slate = (int)__ldcg(p.order + at);
                }
            } else {
                slate = (int)__ldcg(p.order + blockIdx.x);
            }
            sh_next = slate;
        }
        __syncthreads();
        b = sh_next;
        __syncthreads();
    }
    while (b < p.B) {
        const float* sp = p.y_pred + (size_t)b * n;
        const float* yp = p.y_true + (size_t)b * n;
        int cnt_valid = 0;
        unsigned lvl_lo = 0, lvl_hi = 0;                       // counts of labels 1, 2 (16 bits each) and 3, 4
        bool odd = false;
        for (int i = tid; i < n; i += LAMBDA_THREADS) {
            const float y = yp[i];
            const bool pad = (y == p.pad);
            K[i] = pad ? ninf : sp[i];
            cnt_valid += pad ? 0 : 1;
            const int lv = (int)fminf(fmaxf(y, -1.f), 5.f);
            odd = odd || (!pad && !((float)lv == y && lv >= 0 && lv <= 4));
            if (!pad) {
                lvl_lo += (lv == 1) ? 1u : ((lv == 2) ? 0x10000u : 0u);
                lvl_hi += (lv == 3) ? 1u : ((lv == 4) ? 0x10000u : 0u);
            }
        }
        if (tid == 0) sh_m = 0;
        if (tid < 5) sh_lvl[tid] = 0;
        __syncthreads();
        cnt_valid = (int)warp_sum((float)cnt_valid);
        lvl_lo = __reduce_add_sync(LTR_FULL_MASK, lvl_lo);
        lvl_hi = __reduce_add_sync(LTR_FULL_MASK, lvl_hi);
        odd = __any_sync(LTR_FULL_MASK, odd);
        if ((tid & 31) == 0) {
            atomicAdd(&sh_m, cnt_valid);
            atomicAdd(&sh_lvl[1], (int)(lvl_lo & 0xffffu));
            atomicAdd(&sh_lvl[2], (int)(lvl_lo >> 16));
            atomicAdd(&sh_lvl[3], (int)(lvl_hi & 0xffffu));
            atomicAdd(&sh_lvl[4], (int)(lvl_hi >> 16));
            if (odd) atomicOr(&sh_lvl[0], 1);
        }
        __syncthreads();
        const int m = sh_m;
        block_sort_desc<false, LAMBDA_THREADS>(K, I, Y, n, len, TIES_STDSORT, m);
        const int kk = (p.k >= 0 && p.k < n) ? p.k : n;
        const bool hist = !sh_lvl[0];
        // labels 0..4: level L fills the ideal ranks [e_{L+1}, e_L), e_L = #docs with label >= L, so the label at ideal
        // rank r is the number of levels with e_L > r; other labels: ideal order by a bitonic sort of the label values
        const int e4 = sh_lvl[4], e3 = e4 + sh_lvl[3], e2 = e3 + sh_lvl[2], e1 = e2 + sh_lvl[1];
        if (!hist) {
            for (int i = tid; i < len; i += LAMBDA_THREADS) {
                float y = (i < n) ? yp[i] : ninf;
                if (y == p.pad) y = ninf;
                G[i] = y;
            }
            block_bitonic_desc(G, len);
        }
        float part = 0.f;
        for (int r = tid; r < kk; r += LAMBDA_THREADS) {            // ideal DCG@k (lambdaL.py:39)
            const float gain = hist ? (float)((1 << ((e4 > r) + (e3 > r) + (e2 > r) + (e1 > r))) - 1)
                                    : exp2f(fmaxf(G[r], 0.f)) - 1.0f;
            part += gain * invD[r];
        }
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = 0.f;
        part = warp_sum(part);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = part;
        __syncthreads();
        const float maxdcg = fmaxf((red[0] + red[1]) + (red[2] + red[3]), p.eps);
        __syncthreads();
        for (int r = tid; r < n; r += LAMBDA_THREADS) {
            float y = yp[I[r]];
            if (y == p.pad) y = ninf;
            Y[r] = y;                                             // raw label: the mask compares these (:24-27)
            G[r] = (exp2f(fmaxf(y, 0.f)) - 1.0f) / maxdcg;
        }
        __syncthreads();

        // ---- pair loop over the top-k block of real docs -------------------------------------------
        const int lim = kk < m ? kk : m;
        float loss_part = 0.f, cnt_part = 0.f;
        // Fast path: blocks of up to 256 docs whose scaled score range keeps e = exp(-sigma (s_i - s_j)) = 2^-(t_i - c) 2^(t_j - c),
        // t = sigma log2(e) s, inside fp32 as a product of two per-doc factors (K is sorted, NaN first: a NaN / inf range does not
        // qualify).  The tiles leave the eps clamps out and report whether one would have been active (mn, mn2); such a slate,
        // like the long and the wide ones, is done by the exact row-per-thread loop below.
        const float range = (lim > 0) ? (K[0] - K[lim - 1]) * sig_scale : 0.f;
        bool exact = !(lim > 0 && lim <= 256 && fabsf(range) < 126.0f);
        if (!exact) {
            const float mid = 0.5f * K[0] + 0.5f * K[lim - 1];
            for (int i = tid; i < lim; i += LAMBDA_THREADS) {
                const float t = (K[i] - mid) * sig_scale;
                U[i] = ex2_approx(t);
                V[i] = ex2_approx(-t);
            }
            __syncthreads();
            float mn = 0.f, mn2 = 0.f;
            // Can a clamp become active at all?  log2 sig >= -log2(1 + 2^range) > -(range + 1); a pair weight is at most
            // gmax (mu dtab[1] + 1/D_0 - 1/D_{lim-1}), gmax = the gain of the highest label present (histogram labels only).
            bool track = true;
            if (lambda_gain_scheme<SCHEME>() && hist && p.mu >= 0.f && !p.force_track) {
                const int top = e4 ? 4 : (e3 ? 3 : (e2 ? 2 : (e1 ? 1 : 0)));
                const float gmax = (float)((1 << top) - 1) / maxdcg;
                const float dd = invD[0] - invD[lim - 1];
                const float wmax = gmax * (SCHEME == SCHEME_NDCG2 ? dtab[1] : (SCHEME == SCHEME_LAMBDARANK ? dd : fmaf(p.mu, dtab[1], dd)));
                const float lmax = fabsf(range) + 1.0f;
                track = !(lmax < -0.98f * log_eps2 && wmax * lmax < -0.98f * log_eps2);
            }
            if (!track) {
                if (lim > 64) lambda_sym_slate4<SCHEME, false>(U, V, Y, G, invD, dtab, lim, b, p.mu, gpart, loss_part, mn, mn2);
                else lambda_sym_slate<SCHEME, 2, false>(U, V, Y, G, invD, dtab, lim, b, p.mu, gpart, loss_part, mn, mn2);
            } else {
                lambda_slate_tracked<SCHEME>(U, V, Y, G, invD, dtab, lim, b, p.mu, gpart, loss_part, mn, mn2);
            }
            // (a NaN anywhere -- it cannot come from finite inputs here -- also sends the slate to the exact loop)
            exact = __syncthreads_or(!(mn >= log_eps2) || !(mn2 >= log_eps2) || !(loss_part == loss_part)) != 0;
        }
        if (!exact) {
            loss_part *= out_scale;                               // log base: 1 (binary) or ln 2 (natural)
            // number of masked pairs of the block (lambdaL.py:24-30): every ordered pair for ndcgLoss1, otherwise the unordered
            // pairs of docs with different labels = (lim^2 - sum over label values of count^2) / 2
            if (SCHEME == SCHEME_NDCG1) {
                if (tid == 0) cnt_part = (float)lim * (float)lim;
            } else if (hist) {
                if (tid < 5) sh_lvl[tid] = 0;
                __syncthreads();
                unsigned c01 = 0, c23 = 0, c4 = 0;               // per-thread counts (<= 2 docs per thread): 16-bit fields
                for (int i = tid; i < lim; i += LAMBDA_THREADS) {
                    const int lv = (int)Y[i];
                    c01 += (lv == 0) ? 1u : ((lv == 1) ? 0x10000u : 0u);
                    c23 += (lv == 2) ? 1u : ((lv == 3) ? 0x10000u : 0u);
                    c4 += (lv == 4) ? 1u : 0u;
                }
                c01 = __reduce_add_sync(LTR_FULL_MASK, c01);
                c23 = __reduce_add_sync(LTR_FULL_MASK, c23);
                c4 = __reduce_add_sync(LTR_FULL_MASK, c4);
                if ((tid & 31) == 0) {
                    atomicAdd(&sh_lvl[0], (int)(c01 & 0xffffu));
                    atomicAdd(&sh_lvl[1], (int)(c01 >> 16));
                    atomicAdd(&sh_lvl[2], (int)(c23 & 0xffffu));
                    atomicAdd(&sh_lvl[3], (int)(c23 >> 16));
                    atomicAdd(&sh_lvl[4], (int)c4);
                }
                __syncthreads();
                if (tid == 0) {
                    int sq = 0;
                    for (int L = 0; L < 5; ++L) sq += sh_lvl[L] * sh_lvl[L];
                    cnt_part = (float)((lim * lim - sq) / 2);
                }
            } else {
                // labels off the 0..4 grid: count the pairs with equal labels directly (rare)
                int eq = 0;
                for (int i = tid; i < lim; i += LAMBDA_THREADS) {
                    const float yi = Y[i];
                    for (int j = 0; j < i; ++j) eq += (Y[j] == yi) ? 1 : 0;
                }
                cnt_part = -(float)eq;
                if (tid == 0) cnt_part += (float)(lim * (lim - 1) / 2);
            }
            if (p.grad) {
                for (int i = tid; i < n; i += LAMBDA_THREADS) {
                    float g = 0.f;
                    if (i < lim) {
#pragma unroll
                        for (int k8 = 0; k8 < 8; ++k8) g += gpart[k8 * 256 + i];
                    }
                    p.grad[(size_t)b * n + I[i]] = g * grad_scale;     // docs outside the block / padding: 0
                }
            }
        } else {
            loss_part = 0.f;
        for (int i = tid; i < n; i += LAMBDA_THREADS) {
            float g = 0.f;
            if (i < lim) {
                const float yi = Y[i], si = K[i];
                for (int j = 0; j < lim; ++j) {
                    const float yj = Y[j], sj = K[j];
                    float l, dl;
                    if (SCHEME == SCHEME_NDCG1) {
                        // (i, j) as row and (j, i) as column; both always in the mask (labels finite)
                        pair_logprob(si - sj, pair_weight<SCHEME>(i, j, G, invD, dtab, Y, p.mu), sig_scale, log_eps2,
                                     out_scale, l, dl);
                        loss_part += l;
                        cnt_part += 1.0f;
                        g -= dl;
                        pair_logprob(sj - si, pair_weight<SCHEME>(j, i, G, invD, dtab, Y, p.mu), sig_scale, log_eps2,
                                     out_scale, l, dl);
                        g += dl;
                    } else if (yi != yj) {
                        const bool row = yi > yj;                 // this thread's doc is the better-labelled one
                        const float w = pair_weight<SCHEME>(i, j, G, invD, dtab, Y, p.mu);   // symmetric in (i, j)
                        pair_logprob(row ? si - sj : sj - si, w, sig_scale, log_eps2, out_scale, l, dl);
                        if (row) {
                            loss_part += l;
                            cnt_part += 1.0f;
                            g -= dl;
                        } else {
                            g += dl;
                        }
                    }
                }
            }
            if (p.grad) p.grad[(size_t)b * n + I[i]] = (i < m) ? g * grad_scale : 0.f;
        }
        }
        loss_part = warp_sum(loss_part);
        cnt_part = warp_sum(cnt_part);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = loss_part, red[8 + (tid >> 5)] = cnt_part;
        __syncthreads();
        if (tid == 0) {
            p.slate_sum[b] = (red[0] + red[1]) + (red[2] + red[3]);
            p.slate_cnt[b] = (red[8] + red[9]) + (red[10] + red[11]);
            if (p.order) {
                const unsigned idx = p.one_wave ? (unsigned)p.B : gridDim.x + atomicAdd(p.next, 1u);
                sh_next = idx < (unsigned)p.B ? (int)__ldcg(p.order + idx) : p.B;
            } else {
                sh_next = p.next ? (int)(gridDim.x + atomicAdd(p.next, 1u)) : b + (int)gridDim.x;
            }
        }
        __syncthreads();
        b = sh_next;
    }
    if (p.finished) {
        // the last CTA to leave adds up the slates' sums (its first warp: the others are gone at once, nobody waits for the counter)
        __syncthreads();
        if (tid >= 32) return;
        unsigned last = 0u;
        if (tid == 0) {
            __threadfence();
            last = (atomicAdd(p.finished, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        last = __shfl_sync(LTR_FULL_MASK, last, 0);
        if (!last) return;
        __threadfence();
        double ts = 0.0, tc = 0.0;
        for (int i = tid; i < p.B; i += 32) ts += (double)__ldcg(p.slate_sum + i), tc += (double)__ldcg(p.slate_cnt + i);
        ts = warp_sum(ts);
        tc = warp_sum(tc);
        if (tid == 0) {
            p.loss[0] = p.mean ? (float)(-ts / tc) : (float)(-ts);
            if (p.count_out) p.count_out[0] = (float)tc;
            *p.finished = 0u;
            if (p.next) *p.next = 0u;
        }
    }
}

// The schedule of lambda_kernel: slates by decreasing pair work (the length of the top-k block of real docs; the pair loop is
// quadratic in it).  A warp counts one slate's real docs; the last CTA to finish counting-sorts the B keys (512 work classes)
// and clears the claim flags / arrival counters of the one-wave form.  `words` = the schedule's corner of the workspace:
// keys[LAMBDA_ORDER_MAX_B] | order[LAMBDA_ORDER_MAX_B] | claimed[LAMBDA_ORDER_MAX_B] | arrive[LAMBDA_ARRIVE] | cursor -- a region
// other kernels use as scratch too, so everything in it is rewritten per call; `done` is a counter word that is left at 0.
__device__ __forceinline__ void lambda_order_body(const float* __restrict__ y_true, int B, int n, int k, float pad, unsigned* words,
                                                  unsigned* done) {
    __shared__ unsigned hist[LAMBDA_KEYS], start[LAMBDA_KEYS], wtot[8];
    __shared__ int sh_last;
    unsigned* keys = words;
    unsigned* order = words + LAMBDA_ORDER_MAX_B;
    unsigned* claimed = order + LAMBDA_ORDER_MAX_B;
    unsigned* arrive = claimed + LAMBDA_ORDER_MAX_B;
    unsigned* cursor = arrive + LAMBDA_ARRIVE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // The dependent launch (lambda_kernel, a grid that fills every SM) is released at the END of this kernel: released at the
    // top, its CTAs pile onto the SMs while the sort below runs and the pair of launches takes 250 us instead of 50 (measured).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = blockIdx.x * 8 + warp;
    if (b < B) {
        const float* yp = y_true + (size_t)b * n;
        int cnt = 0;
        for (int i = lane; i < n; i += 32) cnt += (__ldg(yp + i) == pad) ? 0 : 1;
        cnt = __reduce_add_sync(LTR_FULL_MASK, cnt);
        const int kk = (k >= 0 && k < n) ? k : n;
        const int lim = cnt < kk ? cnt : kk;
        if (lane == 0) keys[b] = (unsigned)(lim <= 256 ? lim : 256 + (lim >> 3));       // < LAMBDA_KEYS for n <= 2048
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) sh_last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!sh_last) return;                        // (an exited CTA counts as having released the dependent launch)
    __threadfence();
    for (int i = tid; i < LAMBDA_KEYS; i += 256) hist[i] = 0u;
    __syncthreads();
    for (int i = tid; i < B; i += 256) atomicAdd(&hist[__ldcg(keys + i)], 1u);
    __syncthreads();
    // start[c] = number of slates in classes above c: thread t owns classes 511 - 2t and 510 - 2t
    const int c0 = LAMBDA_KEYS - 1 - 2 * tid, c1 = c0 - 1;
    const unsigned h0 = hist[c0], h1 = hist[c1];
    unsigned incl = h0 + h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(LTR_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    unsigned base = incl - (h0 + h1);
    for (int w = 0; w < warp; ++w) base += wtot[w];
    start[c0] = base;
    start[c1] = base + h0;
    __syncthreads();
    for (int i = tid; i < LAMBDA_KEYS; i += 256) hist[i] = 0u;                   // (now the fill counts of the classes)
    __syncthreads();
    for (int i = tid; i < B; i += 256) {
        const unsigned c = __ldcg(keys + i);
        order[start[c] + atomicAdd(&hist[c], 1u)] = (unsigned)i;
        claimed[i] = 0u;
    }
    for (int i = tid; i < LAMBDA_ARRIVE; i += 256) arrive[i] = 0u;
    if (tid == 0) {
        *cursor = 0u;
        *done = 0u;
    }
    asm volatile("griddepcontrol.launch_dependents;");
}

struct LambdaOrderArgs {
    const float* y_true;
    int B, n, k;
    float pad;
    unsigned* words;
    unsigned* done;
};
__global__ void __launch_bounds__(256) lambda_order_kernel_p(const LambdaOrderArgs a) {
    lambda_order_body(a.y_true, a.B, a.n, a.k, a.pad, a.words, a.done);
}

// loss = -sum (reduction "sum") or -sum / count ("mean"); for "mean" the gradient is scaled by 1/count too.
__global__ void __launch_bounds__(256) lambda_finish_kernel(const float* slate_sum, const float* slate_cnt, int B, int mean,
                                                            float* loss, float* count_out, float* grad, long long total,
                                                            unsigned* next) {
    if (next && blockIdx.x == 0 && threadIdx.x == 0) *next = 0u;      // (lambda_kernel is complete: same stream)
    __shared__ double red[2][8];
    __shared__ float sh_inv;
    double s = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) s += (double)slate_sum[i], c += (double)slate_cnt[i];
    s = warp_sum(s);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s, red[1][threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tc = 0.0;
        for (int w = 0; w < 8; ++w) ts += red[0][w], tc += red[1][w];
        if (blockIdx.x == 0) {
            loss[0] = mean ? (float)(-ts / tc) : (float)(-ts);
            if (count_out) count_out[0] = (float)tc;
        }
        sh_inv = tc > 0.0 ? (float)(1.0 / tc) : 0.f;      // no masked pair at all: the mean is NaN, its gradient 0 (as autograd gives)
    }
    __syncthreads();
    if (!mean || !grad) return;
    const float inv = sh_inv;
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) grad[i] *= inv;
}

template <int SCHEME>
static int launch_lambda(const LambdaParams& p, size_t smem, int grid, cudaStream_t st) {
    if (smem > 48 * 1024)
        LTR_CUDA(cudaFuncSetAttribute(lambda_kernel<SCHEME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rc = launch_pdl(lambda_kernel<SCHEME>, grid, LAMBDA_THREADS, smem, st, p);
    return rc ? set_cuda_error(rc, "ltr_lambda_fwdbwd (launch)") : 0;
}

}  // namespace ltr

LTR_API int ltr_lambda_fwdbwd(const float* y_pred, const float* y_true, int B, int n, float eps, float pad_value,
                              int scheme, int k, float sigma, float mu, int reduction_mean, int natural_log,
                              float* slate_sum, float* slate_count, float* loss, float* pair_count, float* grad,
                              void* workspace, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !slate_sum || !slate_count || !loss || B <= 0 || n <= 0)
        return set_error(LTR_EINVAL, "ltr_lambda_fwdbwd: null pointer or non-positive B/n");
    if (scheme < 0 || scheme > 6) return set_error(LTR_EINVAL, "ltr_lambda_fwdbwd: unknown weighing scheme");
    if (n > LAMBDA_MAX_N) return set_error(LTR_ETOOBIG, "ltr_lambda_fwdbwd: slate too long");
    LambdaParams p;
    p.y_pred = y_pred; p.y_true = y_true; p.B = B; p.n = n; p.k = k;
    p.len = 32;
    while (p.len < n) p.len <<= 1;
    p.eps = eps; p.pad = pad_value; p.sigma = sigma; p.mu = mu; p.natural_log = natural_log;
    p.slate_sum = slate_sum; p.slate_cnt = slate_count; p.grad = grad;
    p.next = workspace ? ws_counter(workspace, LTR_WS_LAMBDA_COUNTER) : nullptr;
    p.finished = (workspace && B <= 8192) ? ws_counter(workspace, LTR_WS_LAMBDA_FINISHED) : nullptr;   // (larger batches: lambda_finish_kernel)
    p.loss = loss; p.count_out = pair_count; p.mean = reduction_mean;
    static const bool force_track = getenv("LTR_LAMBDA_TRACK") != nullptr;
    p.force_track = force_track ? 1 : 0;
    const size_t smem = ((size_t)p.len * 6 + 10 * 256) * sizeof(float);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nsm = device_sm_count();
    const int cap = nsm * 8;
    int grid = B < cap ? B : cap;
    int rc = 0;
    // Longest-first schedule.  It costs a launch of ~4 us and pays where the slates' pair work differs and several share an SM
    // (us per call without / with it, valid lengths U{n/2..n}: 600 x 256: 49.8 / 39.0, 1500 x 256: 73.7 / 71.2,
    // 4096 x 256: 171.3 / 164.5, 1024 x 256: 50.7 / 50.3; short slates lose: 2048 x 128: 40.8 / 44.6, 1024 x 100: 20.7 / 25.1)
    p.order = nullptr; p.claimed = nullptr; p.arrive = nullptr; p.cursor = nullptr; p.one_wave = 0; p.nsm = nsm;
    static const bool no_order = getenv("LTR_LAMBDA_NO_ORDER") != nullptr;
    if (workspace && !no_order && n > 128 && B >= 2 * nsm && B <= LAMBDA_ORDER_MAX_B) {
        unsigned* words = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + 1024);
        static_assert(1024 + (3 * LAMBDA_ORDER_MAX_B + LAMBDA_ARRIVE + 2) * 4 <= LTR_WS_RISK_PART1, "schedule overflows its workspace corner");
        const LambdaOrderArgs oa{y_true, B, n, k, pad_value, words, ws_counter(workspace, LTR_WS_LAMBDA_ORDER_DONE)};
        rc = launch_pdl(lambda_order_kernel_p, (B + 7) / 8, 256, 0, st, oa);
        if (rc) return set_cuda_error(rc, "ltr_lambda_fwdbwd (schedule)");
        const int resident = nsm * LAMBDA_MIN_CTAS;
        p.order = words + LAMBDA_ORDER_MAX_B;
        p.claimed = words + 2 * LAMBDA_ORDER_MAX_B;
        p.arrive = words + 3 * LAMBDA_ORDER_MAX_B;
        p.cursor = p.arrive + LAMBDA_ARRIVE;
        p.one_wave = B <= resident ? 1 : 0;
        grid = B < resident ? B : resident;
    }
    switch (scheme) {
        case SCHEME_NONE: rc = launch_lambda<SCHEME_NONE>(p, smem, grid, st); break;
        case SCHEME_NDCG1: rc = launch_lambda<SCHEME_NDCG1>(p, smem, grid, st); break;
        case SCHEME_NDCG2: rc = launch_lambda<SCHEME_NDCG2>(p, smem, grid, st); break;
        case SCHEME_LAMBDARANK: rc = launch_lambda<SCHEME_LAMBDARANK>(p, smem, grid, st); break;
        case SCHEME_NDCG2PP: rc = launch_lambda<SCHEME_NDCG2PP>(p, smem, grid, st); break;
        case SCHEME_GTDIFF: rc = launch_lambda<SCHEME_GTDIFF>(p, smem, grid, st); break;
        default: rc = launch_lambda<SCHEME_GTDIFF_POW>(p, smem, grid, st); break;
    }
    if (rc) return rc;
    if (p.finished && !(reduction_mean && grad)) return 0;         // (the mean's gradient scaling is the second launch)
    const long long total = (long long)B * n;
    int fgrid = reduction_mean && grad ? (int)((total + 256 * 8 - 1) / (256 * 8)) : 1;
    if (fgrid > cap) fgrid = cap;
    if (fgrid < 1) fgrid = 1;
    lambda_finish_kernel<<<fgrid, 256, 0, st>>>(slate_sum, slate_count, B, reduction_mean, loss, pair_count, grad, total, p.next);
    rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_lambda_fwdbwd") : 0;
}
