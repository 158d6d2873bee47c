// K1 fast path: fused forward+backward approxNDCG for slates of at most 128 docs
// (replaces losses/approxNDCG.py:7-53 of the reference; BASELINE configs[1]).
//
// One CTA of 4 warps (one per SM sub-partition) works on one slate at a time; persistent CTAs pull
// slates from an atomic counter and prefetch the next slate's scores/labels into registers.
// The pairwise work exploits  sigma_ji = 1 - sigma_ij  and evaluates every UNORDERED pair once per pass:
//
//   u_i = 2^(t_i - c), v_j = 2^-(t_j - c), t = alpha*log2(e)*s      (2 MUFU per doc)
//   p = u_i * v_j = 2^(t_i - t_j);  r = rcp(1 + p) = sigma_ij;  p*r = sigma_ji      (1 MUFU per pair)
//   w = sigma_ij * sigma_ji   (the sigmoid derivative, exact at both tails)
//
// Pair ownership is a block-circulant: lane a of every warp owns the R = max(2, ceil(m/32)) consecutive
// docs of row block a; at step D it meets column block (a + D) mod T, T = ceil(m/R).  Steps
// D = 1 .. T/2 cover every unordered block pair once, except that for even T the antipodal step
// (D = T/2) meets every pair twice and is weighted 1/2; the R(R-1)/2 pairs inside a lane's own block
// are an extra triangular step.  The T/2 steps are split between the 4 warps (the split rotates
// with the slate index so that no sub-partition always gets the long share).  The column block's
// values and accumulators travel from lane to lane with warp shuffles; row accumulators never
// leave their lane.  Registers stay low (no per-pair state is kept), so 8+ CTAs are resident per
// SM and the MUFU pipe -- one rcp per unordered pair per pass -- is the only saturated resource.
//
//   forward :  pos_i = 1 + sum_{j != i} sigma_ij,  slate value = sum_i G_i / log2(1 + pos_i)
//   backward:  d loss / d s_k = alpha * sum_j w_kj (c_j - c_k),
//              c_i = G_i / (B ln2 (1 + pos_i) log2(1 + pos_i)^2)            
//
// Slates whose scaled score range exceeds the fp32 exponent budget of the factorised form take
// `direct_slate` (one doc per thread, p = ex2(min(alpha*log2e*(s_i - s_j), 126)) per ordered pair).
#pragma once
#include "approx_pair_tiles.cuh"
#include "ltr_common.cuh"

// build-time knobs (A/B experiments: `ltr_b200.build.build_variant`)
#ifndef SYM_MIN_CTAS
#define SYM_MIN_CTAS 8       // resident CTAs per SM the register allocation is tuned for
#endif
#ifndef SYM_CLAIM_AHEAD
#define SYM_CLAIM_AHEAD 0    // 1: claim the next slate index one slate ahead (hides the atomic latency; measured neutral)
#endif
#ifndef SYM_GRID_PER_SM
#define SYM_GRID_PER_SM 0    // CTAs launched per SM (0: as many as are resident)
#endif
#ifndef SYM_PDL
#define SYM_PDL 1            // programmatic dependent launch: overlap this launch's ramp-up with the previous kernel's tail
#endif

namespace ltr {

// discount prefix sums P[r] = sum_{q<=r} 1/log2(1+q), r = 0..128, for the histogram ideal DCG of the fast path
// (summed in fp64 and rounded once by ltr_fill_disc_prefix; a uniform-index constant-bank read, no per-CTA table
// load).  The reference sums these terms in fp32; four rounded table entries are well inside its own error.
__constant__ float c_disc_prefix[129];

struct ApproxParams {
    const float* s;
    const float* y;
    int B, n, npad;
    float alpha, eps, pad, inv_B;
    const double* disc_prefix;
    int disc_len;
    float* slate_out;
    float* loss_out;
    float* grad;
    unsigned* ws;   // [0] next-slate counter, [2..3] 64-bit word: non-finite flag << 63 | finished-CTA count << 52 |
                    // fixed-point loss sum (all left at 0)
    int clamp;      // 1: apply the eps clamp per pair exactly (large eps), 0: it is below fp32 resolution
    float* cbuf;    // split path: [B, n] d loss / d pos of every compacted doc (rank launch -> gradient launch)
    float* part;    // split path: [B, chunks] partial slate values
    int rows_per_cta;
};

constexpr int SYM_N = 128;        // longest slate of the fast path
constexpr int SYM_THREADS = 128;  // 4 warps: one per SM sub-partition
constexpr int SYM_MAX_LABEL = 4;  // integer labels 0..4 take the histogram ideal-DCG (others: sort)

struct SymShared {
    float u[SYM_N];               // 2^(t - c)   (direct form: raw score)
    float v[SYM_N];               // 2^-(t - c)
    float c[SYM_N];               // d loss / d pos_i
    float part[4][SYM_N];         // per-warp partial sums of every doc (ranks, then gradients): row side and column side combined
    float sortbuf[SYM_N];
    int4 wstat[4];                // per warp: {valid count | bad-label flag << 16, min key, max key, packed label counts}
    float red[4];
    int next_b;
};

// Everything a thread keeps about its own doc across the phases of a slate.
struct SymDoc {
    bool valid;
    int dst;        // compacted position of this thread's doc
    float gain;     // normalised gain G_i
};

// Shared middle of a slate: soft rank -> slate value and d loss / d pos_i (sh.c), then fetch the next
// slate index and prefetch its scores/labels.  `sum` = sum_{j != i} sigma_ij of this thread's doc.
// Per-CTA running state that survives from slate to slate (meaningful in thread 0 only).
struct SymCarry {
    int claimed;          // slate index already claimed from the global counter for the NEXT iteration
    long long fx;         // sum of this CTA's slate values in 2^-40 fixed point (order-independent, so the
                          // batch loss is bitwise reproducible although slates are handed out dynamically)
    int nonfinite;
};
constexpr float kFixScale = 67108864.0f;        // 2^26: slate values (a few units at most) of up to 2^22 slates fit 52 bits
constexpr int kFixCountShift = 52;               // the finished-CTA count rides in bits 52..62 of the same word

__device__ __forceinline__ void carry_add(SymCarry& cy, float v) {
    if (fabsf(v) < 4.0e6f) cy.fx += __float2ll_rn(v * kFixScale);
    else cy.nonfinite = 1;                       // NaN / inf / absurd: the batch loss becomes NaN
}

__device__ __forceinline__ void slate_middle(const ApproxParams& p, SymShared& sh, int b, const SymDoc& doc, float sum,
                                             SymCarry& cy, float& next_s, float& next_y, int& next_b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float ndcg_part = 0.f;
    if (doc.valid) {
        const float pos1 = 2.0f + sum;                       // 1 + pos_i
        const float l2 = lg2_approx(pos1);
        const float il2 = rcp_approx(l2);
        ndcg_part = doc.gain * il2;
        sh.c[doc.dst] = (p.inv_B * (1.0f / kLn2)) * doc.gain * il2 * il2 * rcp_approx(pos1);
    }
    ndcg_part = warp_sum(ndcg_part);
    if (lane == 0) sh.red[warp] = ndcg_part;
    if (tid == 0) {
#if SYM_CLAIM_AHEAD
        sh.next_b = cy.claimed;                                   // claimed one slate ago: latency long hidden
        cy.claimed = (int)(gridDim.x + atomicAdd(&p.ws[0], 1u));  // result is first read a whole slate later
#else
        sh.next_b = (int)(gridDim.x + atomicAdd(&p.ws[0], 1u));
#endif
    }
    __syncthreads();
    if (tid == 0) {
        const float val = (sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]);
        p.slate_out[b] = val;
        carry_add(cy, val);
    }
    next_b = sh.next_b;
    if (next_b < p.B && tid < p.n) {
        next_s = ld_stream(p.s + (size_t)next_b * p.n + tid);
        next_y = ld_stream(p.y + (size_t)next_b * p.n + tid);
    }
}

// Slates whose scaled score range does not fit the factorised form (rare): one doc per thread,
// p = ex2(min(alpha*log2e*(s_i - s_j), 126)) per ordered pair, recomputed in the gradient pass.
// `raw` holds the raw scores of the m compacted docs.  The two loops are out-of-line helpers with by-value
// arguments only (a by-reference argument of a __noinline__ function would pin the caller's slate state in
// local memory on the fast path as well).
__device__ __noinline__ float direct_rank_sum(const float* raw, int m, float scale, float si) {
    float sum = -0.5f;                                           // the j == i term below is exactly 1/2
    for (int j = 0; j < m; ++j) {
        const float pp = ex2_capped((si - raw[j]) * scale);
        sum += rcp_approx(pp + 1.0f);
    }
    return sum;
}
__device__ __noinline__ float direct_grad_sum(const float* raw, const float* c, int m, float scale, float si, float ci) {
    float acc = 0.f;
    for (int j = 0; j < m; ++j) {
        const float pp = ex2_capped((si - raw[j]) * scale);
        const float rr = rcp_approx(pp + 1.0f);
        acc = fmaf(rr * (pp * rr), c[j] - ci, acc);
    }
    return acc;
}
__device__ __forceinline__ void direct_slate(const ApproxParams& p, SymShared& sh, int b, int m, float scale,
                                             const SymDoc& doc, SymCarry& cy, float& next_s, float& next_y, int& next_b) {
    const int tid = threadIdx.x;
    const float si = doc.valid ? sh.u[doc.dst] : 0.f;
    float sum = 0.f;
    if (doc.valid) sum = direct_rank_sum(sh.u, m, scale, si);
    slate_middle(p, sh, b, doc, sum, cy, next_s, next_y, next_b);
    if (p.grad != nullptr && tid < p.n) {
        float acc = 0.f;
        if (doc.valid) acc = direct_grad_sum(sh.u, sh.c, m, scale, si, sh.c[doc.dst]);
        p.grad[(size_t)b * p.n + tid] = p.alpha * acc;
    }
    __syncthreads();
}

// One slate in the factorised form, R docs per lane.  Called by all 128 threads with CTA-uniform arguments.
template <int R>
__device__ __forceinline__ void sym_slate(const ApproxParams& p, SymShared& sh, int b, int m, const SymDoc& doc,
                                          SymCarry& cy, float& next_s, float& next_y, int& next_b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = (m + R - 1) / R;            // active lanes = row blocks
    const int S = T >> 1;                     // circulant steps D = 1 .. T/2 (D = 0 is the in-block triangle)
    const int role = (warp + b) & 3;
    const int d_lo = 1 + ((role * S) >> 2);
    const int d_hi = 1 + (((role + 1) * S) >> 2);
    const int steps = d_hi - d_lo;
    // for even T the antipodal step D = T/2 meets every pair twice and is weighted 1/2: it is the last step of role 3
    const bool anti = ((T & 1) == 0) && (role == 3) && (steps > 0);
    const int full = steps - (anti ? 1 : 0);
    const bool active = lane < T;
    const int src = active ? ((lane + 1 == T) ? 0 : lane + 1) : lane;
    int cb = lane + d_lo;
    if (cb >= T) cb -= T;
    if (!active) cb = 0;
    int ret = lane - (d_hi - 1);              // the lane that ends up holding this lane's column accumulators
    if (ret < 0) ret += T;
    if (steps == 0 || !active) ret = lane;    // nothing accumulated: the column sums are zero

    float urow[R], vcol[R], acc_r[R], acc_r_hi[R], acc_c[R];
    lds_vec<R>(sh.u + lane * R, urow);
    lds_vec<R>(sh.v + cb * R, vcol);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc_r[r] = 0.f;
        acc_r_hi[r] = 0.f;
        acc_c[r] = 0.f;
    }

    // ---------------- rank pass: one MUFU.RCP per unordered pair -----------------------------
    if (R > 1 && role == 0) {                 // pairs inside the lane's own block
        float vrow[R];
        lds_vec<R>(sh.v + lane * R, vrow);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = r + 1; c < R; ++c) {
                const float pp = urow[r] * vrow[c];
                const float rr = rcp_approx(pp + 1.0f);
                acc_r[r] += rr;
                acc_r[c] = fmaf(pp, rr, acc_r[c]);
            }
        }
    }
#pragma unroll 1
    for (int i = 0; i < full; ++i) {
        if (i) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
            }
        }
        rank_tile<R>(urow, vcol, 1.0f, acc_r, acc_r_hi, acc_c);
    }
    if (anti) {
        if (full) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
            }
        }
        rank_tile<R>(urow, vcol, 0.5f, acc_r, acc_r_hi, acc_c);
    }
    // the column accumulators go home; a doc's partial is (sum sigma_ij as a row) - (sum sigma_ij as a column)
    {
        float out[R];
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = (acc_r[r] + acc_r_hi[r]) - __shfl_sync(LTR_FULL_MASK, acc_c[r], ret);
        if (active) sts_vec<R>(sh.part[warp] + lane * R, out);
    }
    __syncthreads();

    // ---------------- per doc: soft rank, slate value, d loss / d pos --------------------------
    float sum = 0.f;
    if (doc.valid) {
        const float parts = (sh.part[0][doc.dst] + sh.part[1][doc.dst]) + (sh.part[2][doc.dst] + sh.part[3][doc.dst]);
        // column side: sum sigma_ji = (#rows met, antipodal ones at 1/2) - sum sigma_ij = R (T-1)/2 - cols.
        // padded tail of the last block (u = v = 0): as a column it reads sigma_ij = 1 on the row side.
        const int npadc = T * R - m;
        const int e = T - 1 - doc.dst / R;               // circulant distance to the last block
        const float f = (2 * e == T) ? 0.5f : ((2 * e < T) ? 1.0f : 0.0f);
        sum = (parts - f * (float)npadc) + 0.5f * (float)(R * (T - 1));
    }
    slate_middle(p, sh, b, doc, sum, cy, next_s, next_y, next_b);
    if (p.grad == nullptr) {
        __syncthreads();
        return;
    }

    // ---------------- gradient pass: sigma recomputed, one MUFU.RCP per unordered pair ---------
    float crow[R], ccol[R];
    lds_vec<R>(sh.c + lane * R, crow);
    lds_vec<R>(sh.c + cb * R, ccol);
    lds_vec<R>(sh.v + cb * R, vcol);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc_r[r] = 0.f;
        acc_r_hi[r] = 0.f;
        acc_c[r] = 0.f;
    }
    if (R > 1 && role == 0) {
        float vrow[R];
        lds_vec<R>(sh.v + lane * R, vrow);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = r + 1; c < R; ++c) {
                const float pp = urow[r] * vrow[c];
                const float rr = rcp_approx(pp + 1.0f);
                const float e = (rr * (pp * rr)) * (crow[c] - crow[r]);
                acc_r[r] += e;
                acc_r[c] -= e;
            }
        }
    }
#pragma unroll 1
    for (int i = 0; i < full; ++i) {
        if (i) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                ccol[r] = __shfl_sync(LTR_FULL_MASK, ccol[r], src);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
            }
        }
        grad_tile<R>(urow, vcol, crow, ccol, acc_r, acc_r_hi, acc_c);
    }
    if (anti) {                                 // weight 1/2 goes onto the c differences
        if (full) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                ccol[r] = __shfl_sync(LTR_FULL_MASK, ccol[r], src);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
            }
        }
        float ch[R], cch[R];
#pragma unroll
        for (int r = 0; r < R; ++r) ch[r] = 0.5f * crow[r], cch[r] = 0.5f * ccol[r];
        grad_tile<R>(urow, vcol, ch, cch, acc_r, acc_r_hi, acc_c);
    }
    {
        float out[R];
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = (acc_r[r] + acc_r_hi[r]) + __shfl_sync(LTR_FULL_MASK, acc_c[r], ret);
        if (active) sts_vec<R>(sh.part[warp] + lane * R, out);
    }
    __syncthreads();
    if (tid < p.n) {
        float sum2 = 0.f;
        if (doc.valid) sum2 = (sh.part[0][doc.dst] + sh.part[1][doc.dst]) + (sh.part[2][doc.dst] + sh.part[3][doc.dst]);
        p.grad[(size_t)b * p.n + tid] = p.alpha * sum2;          // padded docs: exactly 0
    }
}

__global__ void __launch_bounds__(SYM_THREADS, SYM_MIN_CTAS) approx_ndcg_sym_kernel(const ApproxParams p) {
    __shared__ __align__(16) SymShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const float scale = p.alpha * kLog2e;

#if SYM_PDL
    // let the next launch on the stream start its ramp-up now; wait for everything before us to be complete and
    // visible before touching global memory (inputs may come from the previous kernel, ws from the previous call)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    int b = (int)blockIdx.x;                      // first slate is static, the rest come from the counter
    float cur_s = 0.f, cur_y = p.pad;
    if (b < p.B && tid < n) {
        cur_s = ld_stream(p.s + (size_t)b * n + tid);
        cur_y = ld_stream(p.y + (size_t)b * n + tid);
    }
    SymCarry cy;
    cy.fx = 0;
    cy.nonfinite = 0;
    cy.claimed = 0;
#if SYM_CLAIM_AHEAD
    if (tid == 0) cy.claimed = (int)(gridDim.x + atomicAdd(&p.ws[0], 1u));
#endif

    while (b < p.B) {
        // ---- compaction offsets + per-slate statistics: 5 warp collectives, one barrier ---------
        const float sv = cur_s, yv = cur_y;
        SymDoc doc;
        doc.valid = (tid < n) && (yv != p.pad);
        const float yc = fmaxf(yv, 0.f);                         // approxNDCG.py:37-38 clamp
        const int lvl = (int)fminf(yc, 100.f);
        const bool small_int = (yc == (float)lvl) && (lvl <= SYM_MAX_LABEL);
        const unsigned bal = __ballot_sync(LTR_FULL_MASK, doc.valid);
        const unsigned bad = __ballot_sync(LTR_FULL_MASK, doc.valid && !small_int);
        const int key = float_order_key(sv);
        const int kmin = __reduce_min_sync(LTR_FULL_MASK, doc.valid ? key : 0x7fffffff);
        const int kmax = __reduce_max_sync(LTR_FULL_MASK, doc.valid ? key : (int)0x80000000);
        // label histogram: one 8-bit field per level 1..4 (<= 32 per warp, <= 128 per slate)
        const unsigned field = (doc.valid && small_int && lvl >= 1) ? (1u << (8 * (lvl - 1))) : 0u;
        const unsigned cnts = __reduce_add_sync(LTR_FULL_MASK, field);
        if (lane == 0) sh.wstat[warp] = make_int4(__popc(bal) | (bad ? 0x10000 : 0), kmin, kmax, (int)cnts);
        __syncthreads();
        int m = 0, off = 0, any_bad = 0;
        int gmin = 0x7fffffff, gmax = (int)0x80000000;
        unsigned packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int4 st = sh.wstat[k];
            const int cnt = st.x & 0xffff;
            off += (k < warp) ? cnt : 0;
            m += cnt;
            any_bad |= st.x >> 16;
            gmin = min(gmin, st.y);
            gmax = max(gmax, st.z);
            packed += (unsigned)st.w;
        }
        doc.dst = off + __popc(bal & ((1u << lane) - 1u));
        doc.gain = 0.f;

        int next_b = 0;
        float next_s = 0.f, next_y = p.pad;
        if (m == 0) {                                            // fully padded slate: value 0, no gradient
            if (tid == 0) {
                p.slate_out[b] = 0.f;
#if SYM_CLAIM_AHEAD
                sh.next_b = cy.claimed;
                cy.claimed = (int)(gridDim.x + atomicAdd(&p.ws[0], 1u));
#else
                sh.next_b = (int)(gridDim.x + atomicAdd(&p.ws[0], 1u));
#endif
            }
            if (tid < n && p.grad != nullptr) p.grad[(size_t)b * n + tid] = 0.f;
            __syncthreads();
            next_b = sh.next_b;
            if (next_b < p.B && tid < n) {
                next_s = ld_stream(p.s + (size_t)next_b * n + tid);
                next_y = ld_stream(p.y + (size_t)next_b * n + tid);
            }
            __syncthreads();
        } else {
            // ---- ideal DCG over ALL positions (approxNDCG.py:41-43) -----------------------------
            float maxdcg, gain;
            if (!any_bad) {
                // labels 0..4: level L fills ideal ranks (e_{L+1}, e_L], e_L = #docs with label >= L, so
                // maxDCG = sum_L (2^L - 1)(P[e_L] - P[e_{L+1}]) = 8 P[e4] + 4 P[e3] + 2 P[e2] + P[e1]
                const int e4 = packed >> 24;
                const int e3 = e4 + ((packed >> 16) & 255);
                const int e2 = e3 + ((packed >> 8) & 255);
                const int e1 = e2 + (packed & 255);
                const float dcg = fmaf(8.0f, c_disc_prefix[e4], fmaf(4.0f, c_disc_prefix[e3], fmaf(2.0f, c_disc_prefix[e2], c_disc_prefix[e1])));
                maxdcg = fmaxf(dcg, p.eps);
                gain = (float)((1 << lvl) - 1);
            } else {
                // general labels: bitonic sort (descending) of the clamped labels
                sh.sortbuf[tid] = doc.valid ? yc : -1.0f;        // -1 sorts after every clamped label
                block_bitonic_desc(sh.sortbuf, SYM_N);
                const float lab = sh.sortbuf[tid];
                float part = (lab >= 0.f) ? gain_exp2(lab) / log2f(2.0f + (float)tid) : 0.f;
                part = warp_sum(part);
                if (lane == 0) sh.red[warp] = part;
                __syncthreads();
                maxdcg = fmaxf((sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]), p.eps);
                gain = gain_exp2(yc);
                __syncthreads();
            }
            doc.gain = __fdividef(gain, maxdcg);                 // G_i (approxNDCG.py:44)

            // ---- compacted per-doc quantities -----------------------------------------------------
            const float smin = float_from_key(gmin), smax = float_from_key(gmax);
            const float range = (smax - smin) * scale;
            const bool fact = fabsf(range) < 126.0f;             // false for inf / NaN ranges too
            const float smid = 0.5f * smin + 0.5f * smax;
            if (doc.valid) {
                if (fact) {
                    const float tc = (sv - smid) * scale;
                    sh.u[doc.dst] = ex2_approx(tc);
                    sh.v[doc.dst] = ex2_approx(-tc);
                } else {
                    sh.u[doc.dst] = sv;
                }
            }
            if (tid >= m) {                                      // zero tail: p = 0, w = 0 for padding
                sh.u[tid] = 0.f;
                sh.v[tid] = 0.f;
                sh.c[tid] = 0.f;
            }
            __syncthreads();

            if (!fact) direct_slate(p, sh, b, m, scale, doc, cy, next_s, next_y, next_b);
            else if (m > 96) sym_slate<4>(p, sh, b, m, doc, cy, next_s, next_y, next_b);
            else if (m > 64) sym_slate<3>(p, sh, b, m, doc, cy, next_s, next_y, next_b);
            else sym_slate<2>(p, sh, b, m, doc, cy, next_s, next_y, next_b);   // (short slates leave lanes idle)
        }
        b = next_b;
        cur_s = next_s;
        cur_y = next_y;
    }

    // ---- batch loss: ONE 64-bit atomic per CTA carries its fixed-point sum (low 52 bits) and counts it as finished
    //      (bits 52..62); bit 63 is a sticky "a slate value was not finite" flag.  The CTA that completes the count
    //      already holds the total and the flag in the value its own atomic returned: no fence, no second round trip ----
    if (tid == 0) {
        unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(p.ws + 2);
        if (cy.nonfinite) atomicOr(acc64, 1ull << 63);               // same address, same thread: ordered before the add
        const unsigned long long mine = (unsigned long long)cy.fx + (1ull << kFixCountShift);
        const unsigned long long after = atomicAdd(acc64, mine) + mine;
        if (((after >> kFixCountShift) & 0x7ffull) == (unsigned long long)gridDim.x) {
            const double total = (double)(after & ((1ull << kFixCountShift) - 1ull));
            const double mean = total / ((double)kFixScale * (double)p.B);
            p.loss_out[0] = (after >> 63) ? __int_as_float(0x7fc00000) : (float)(-mean);       // approxNDCG.py:53
            p.ws[0] = 0u;
            *acc64 = 0ull;
        }
    }
}

}  // namespace ltr
