// K1 fast path, second generation: fused forward+backward approxNDCG for slates of at most 128 docs with ONE WARP
// PER SLATE (replaces losses/approxNDCG.py:7-53 of the reference; BASELINE configs[1]).
//
// The 4-warps-per-slate kernel (approx_ndcg_sym.cuh) spends two thirds of its issue slots on per-slate work that
// every warp repeats (statistics, compaction offsets, partial-sum exchange through shared memory, five CTA
// barriers).  Here a warp owns the whole slate: no CTA barrier, no cross-warp exchange, row accumulators AND
// column accumulators stay in registers (the column block comes home with one shuffle), per-slate statistics are
// computed once.  The pair arithmetic is the same block-circulant schedule:
//
//   u_i = 2^(t_i - c), v_j = 2^-(t_j - c), t = alpha*log2(e)*s      (2 MUFU per doc)
//   p = u_i * v_j;  r = rcp(1 + p) = sigma_ij;  p*r = sigma_ji       (1 MUFU per UNORDERED pair per pass)
//
// Lane a owns the R = 2/3/4 consecutive compacted docs of row block a and meets column block (a + D) mod T at
// step D = 1 .. T/2, T = ceil(m/R) <= 32 (for even T the antipodal step meets every pair twice and is weighted
// 1/2); the R(R-1)/2 pairs inside the lane's own block are a triangular extra step.
//
// Docs are held interleaved (lane l holds docs l, l+32, l+64, l+96: coalesced scalar loads/stores for any n);
// the loss is permutation invariant, so compaction may use any order.  Warps pull slates from an
// atomic counter one slate ahead and prefetch the next slate's row into registers.
#pragma once
#include "approx_ndcg_sym.cuh"

#ifndef WPS_MIN_CTAS
#define WPS_MIN_CTAS 4       // resident CTAs (of 4 independent warps) per SM
#endif

namespace ltr {

constexpr int WPS_THREADS = 128;
constexpr int WPS_WARPS = WPS_THREADS / 32;
constexpr int WPS_SLOTS = 136;    // 128 docs + the zero tail of the last row block

struct WarpShared {
    float u[WPS_SLOTS];      // 2^(t - c)   (direct form: raw score)
    float v[WPS_SLOTS];      // 2^-(t - c)
    float g[WPS_SLOTS];      // normalised gain G_i of the compacted doc (direct form: later d loss / d pos)
    float out[WPS_SLOTS];    // gradient of the compacted doc (general labels: clamped labels while ranking them)
};

struct WpsCarry {
    long long fx;            // sum of this warp's slate values in fixed point (order independent -> reproducible)
    int nonfinite;
};
constexpr float kWpsFixScale = 67108864.0f;      // 2^26
__device__ __forceinline__ void wps_carry_add(WpsCarry& cy, float v) {
    if (fabsf(v) < 4.0e6f) cy.fx += __float2ll_rn(v * kWpsFixScale);
    else cy.nonfinite = 1;                       // NaN / inf / absurd: the batch loss becomes NaN
}

// ideal rank of a label among the m compacted labels (general labels only; ties are value-identical, any order)
__device__ __noinline__ float wps_ideal_term(const float* labs, int m, float y, int dst) {
    int rank = 0;
    for (int k = 0; k < m; ++k) {
        const float o = labs[k];
        rank += (o > y || (o == y && k < dst)) ? 1 : 0;
    }
    return gain_exp2(y) / log2f(2.0f + (float)rank);
}

// One slate in the factorised form, R docs per lane, all 32 lanes.  `vmask`/`dst4`: validity bits and compacted
// positions (one byte each) of the lane's four interleaved docs; the gradient is written by the caller from sh.out.
template <int R>
__device__ __forceinline__ float wps_slate(const ApproxParams& p, WarpShared& sh, int m, int lane) {
    const int T = (m + R - 1) / R;            // active lanes = row blocks
    const int S = T >> 1;                     // circulant steps D = 1 .. T/2
    const bool even = (T & 1) == 0;
    const bool active = lane < T;
    const int nxt = active ? ((lane + 1 == T) ? 0 : lane + 1) : lane;
    int ret = lane - S;                       // the lane that ends up holding this lane's column accumulators
    if (ret < 0) ret += T;
    if (!active) ret = lane;

    float urow[R], vrow[R], grow[R], vcol[R], acc_r[R], acc_r_hi[R], acc_c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int d = lane * R + r;
        urow[r] = active ? sh.u[d] : 0.f;     // inactive lanes: p = 0, finite garbage that nobody reads
        vrow[r] = active ? sh.v[d] : 0.f;
        grow[r] = (d < m) ? sh.g[d] : 0.f;
        vcol[r] = sh.v[nxt * R + r];
        acc_r[r] = 0.f;
        acc_r_hi[r] = 0.f;
        acc_c[r] = 0.f;
    }

    // ---------------- rank pass: one MUFU.RCP per unordered pair -----------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {             // pairs inside the lane's own block
#pragma unroll
        for (int c = r + 1; c < R; ++c) {
            const float pp = urow[r] * vrow[c];
            const float rr = rcp_approx(pp + 1.0f);
            acc_r[r] += rr;
            acc_r[c] = fmaf(pp, rr, acc_r[c]);
        }
    }
#pragma unroll 1
    for (int D = 1; D <= S; ++D) {
        rank_tile<R>(urow, vcol, (even && D == S) ? 0.5f : 1.0f, acc_r, acc_r_hi, acc_c);   // antipodal: pairs met twice
        if (D < S) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], nxt);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], nxt);
            }
        }
    }

    // ---------------- per doc: soft rank, slate value, d loss / d pos --------------------------
    // column side: sum sigma_ji = (#rows met, antipodal ones at 1/2) - sum sigma_ij = R (T-1)/2 - cols.
    // padded tail of the last block (u = v = 0): as a column it reads sigma_ij = 1 on the row side.
    const int npadc = T * R - m;
    const int e = T - 1 - lane;               // circulant distance to the last block
    const float f = (2 * e == T) ? 0.5f : ((2 * e < T) ? 1.0f : 0.0f);
    const float rowfix = f * (float)npadc;
    const float colcnt = 0.5f * (float)(R * (T - 1));
    const float kc = p.inv_B * (1.0f / kLn2);
    float crow[R], ccol[R];
    float part = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float cols = __shfl_sync(LTR_FULL_MASK, acc_c[r], ret);
        const float sum = ((acc_r[r] + acc_r_hi[r]) - rowfix) + (colcnt - cols);
        const float pos1 = 2.0f + sum;                       // 1 + pos_i
        const float il2 = rcp_approx(lg2_approx(pos1));
        const float gi = grow[r] * il2;                      // grow = 0 beyond the slate's docs
        const bool real = lane * R + r < m;
        part += real ? gi : 0.f;
        crow[r] = real ? kc * gi * il2 * rcp_approx(pos1) : 0.f;
    }
    part = warp_sum(part);
    if (p.grad == nullptr) return part;

    // ---------------- gradient pass: sigma recomputed, one MUFU.RCP per unordered pair ---------
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ccol[r] = __shfl_sync(LTR_FULL_MASK, crow[r], nxt);
        vcol[r] = sh.v[nxt * R + r];
        acc_r[r] = 0.f;
        acc_r_hi[r] = 0.f;
        acc_c[r] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int c = r + 1; c < R; ++c) {
            const float pp = urow[r] * vrow[c];
            const float rr = rcp_approx(pp + 1.0f);
            const float ee = (rr * (pp * rr)) * (crow[c] - crow[r]);
            acc_r[r] += ee;
            acc_r[c] -= ee;
        }
    }
#pragma unroll 1
    for (int D = 1; D <= S; ++D) {
        if (even && D == S) {                   // antipodal step: weight 1/2 goes onto the c differences
            float ch[R], cch[R];
#pragma unroll
            for (int r = 0; r < R; ++r) ch[r] = 0.5f * crow[r], cch[r] = 0.5f * ccol[r];
            grad_tile<R>(urow, vcol, ch, cch, acc_r, acc_r_hi, acc_c);
        } else {
            grad_tile<R>(urow, vcol, crow, ccol, acc_r, acc_r_hi, acc_c);
        }
        if (D < S) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], nxt);
                ccol[r] = __shfl_sync(LTR_FULL_MASK, ccol[r], nxt);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], nxt);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float cols = __shfl_sync(LTR_FULL_MASK, acc_c[r], ret);
        if (active) sh.out[lane * R + r] = p.alpha * ((acc_r[r] + acc_r_hi[r]) + cols);
    }
    return part;
}

__global__ void __launch_bounds__(WPS_THREADS, WPS_MIN_CTAS) approx_ndcg_warp_kernel(const ApproxParams p) {
    __shared__ __align__(16) WarpShared shs[WPS_WARPS];
    __shared__ long long cta_fx[WPS_WARPS];
    __shared__ int cta_bad[WPS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpShared& sh = shs[warp];
    const int n = p.n;
    const float scale = p.alpha * kLog2e;
    const int nwarps = (int)gridDim.x * WPS_WARPS;
    const unsigned ltmask = (1u << lane) - 1u;

#if SYM_PDL
    // let the next launch on the stream start its ramp-up now; wait for everything before us to be complete and
    // visible before touching global memory (inputs may come from the previous kernel, ws from the previous call)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    int b = (int)blockIdx.x * WPS_WARPS + warp;   // first slate is static, the rest come from the counter
    int claimed = 0;                              // (lane 0) the slate after this one, claimed one slate ahead
    if (lane == 0) claimed = nwarps + (int)atomicAdd(&p.ws[0], 1u);
    float cur_s[4], cur_y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cur_s[j] = 0.f;
        cur_y[j] = p.pad;
        const int doc = 32 * j + lane;
        if (b < p.B && doc < n) {
            cur_s[j] = ld_stream(p.s + (size_t)b * n + doc);
            cur_y[j] = ld_stream(p.y + (size_t)b * n + doc);
        }
    }
    WpsCarry cy;
    cy.fx = 0;
    cy.nonfinite = 0;

    while (b < p.B) {
        // ---- next slate: index claimed during the previous slate, row prefetched during this one ----
        const int next_b = __shfl_sync(LTR_FULL_MASK, claimed, 0);
        if (lane == 0 && next_b < p.B) claimed = nwarps + (int)atomicAdd(&p.ws[0], 1u);
        float next_s[4], next_y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            next_s[j] = 0.f;
            next_y[j] = p.pad;
            const int doc = 32 * j + lane;
            if (next_b < p.B && doc < n) {
                next_s[j] = ld_stream(p.s + (size_t)next_b * n + doc);
                next_y[j] = ld_stream(p.y + (size_t)next_b * n + doc);
            }
        }

        // ---- validity, compaction offsets, per-slate statistics (once per slate) -------------------
        bool valid[4];
        int dst[4], lvl[4];
        float yc[4];
        int m = 0, kmin = 0x7fffffff, kmax = (int)0x80000000;
        unsigned field = 0;
        bool odd_label = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            valid[j] = (32 * j + lane < n) && (cur_y[j] != p.pad);
            const unsigned bal = __ballot_sync(LTR_FULL_MASK, valid[j]);
            dst[j] = m + __popc(bal & ltmask);
            m += __popc(bal);
            yc[j] = fmaxf(cur_y[j], 0.f);                        // approxNDCG.py:37-38 clamp
            lvl[j] = (int)fminf(yc[j], 100.f);
            const bool small_int = (yc[j] == (float)lvl[j]) && (lvl[j] <= SYM_MAX_LABEL);
            odd_label |= valid[j] && !small_int;
            // label histogram: one 8-bit field per level 1..4 (<= 128 per slate)
            field += (valid[j] && small_int && lvl[j] >= 1) ? (1u << (8 * (lvl[j] - 1))) : 0u;
            const int key = float_order_key(cur_s[j]);
            kmin = min(kmin, valid[j] ? key : 0x7fffffff);
            kmax = max(kmax, valid[j] ? key : (int)0x80000000);
        }
        if (m == 0) {                                            // fully padded slate: value 0, no gradient
            if (lane == 0) p.slate_out[b] = 0.f;
            if (p.grad != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (32 * j + lane < n) p.grad[(size_t)b * n + 32 * j + lane] = 0.f;
            }
        } else {
            const bool any_odd = __any_sync(LTR_FULL_MASK, odd_label);
            kmin = __reduce_min_sync(LTR_FULL_MASK, kmin);
            kmax = __reduce_max_sync(LTR_FULL_MASK, kmax);

            // ---- ideal DCG over ALL positions (approxNDCG.py:41-43) -----------------------------
            float maxdcg;
            if (!any_odd) {
                // labels 0..4: level L fills ideal ranks (e_{L+1}, e_L], e_L = #docs with label >= L, so
                // maxDCG = sum_L (2^L - 1)(P[e_L] - P[e_{L+1}]) = 8 P[e4] + 4 P[e3] + 2 P[e2] + P[e1]
                const unsigned packed = __reduce_add_sync(LTR_FULL_MASK, field);
                const int e4 = packed >> 24;
                const int e3 = e4 + ((packed >> 16) & 255);
                const int e2 = e3 + ((packed >> 8) & 255);
                const int e1 = e2 + (packed & 255);
                const float dcg = fmaf(8.0f, c_disc_prefix[e4], fmaf(4.0f, c_disc_prefix[e3], fmaf(2.0f, c_disc_prefix[e2], c_disc_prefix[e1])));
                maxdcg = fmaxf(dcg, p.eps);
            } else {
                // general labels: every doc finds its ideal rank by counting (rare path)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (valid[j]) sh.out[dst[j]] = yc[j];
                __syncwarp();
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (valid[j]) part += wps_ideal_term(sh.out, m, yc[j], dst[j]);
                maxdcg = fmaxf(warp_sum(part), p.eps);
                __syncwarp();
            }
            float gain[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float raw = any_odd ? gain_exp2(yc[j]) : (float)((1 << lvl[j]) - 1);
                gain[j] = __fdividef(raw, maxdcg);               // G_i (approxNDCG.py:44)
            }

            // ---- compacted per-doc quantities -----------------------------------------------------
            const float smin = float_from_key(kmin), smax = float_from_key(kmax);
            const float range = (smax - smin) * scale;
            const bool fact = fabsf(range) < 126.0f;             // false for inf / NaN ranges too
            const float smid = 0.5f * smin + 0.5f * smax;
            float val;
            if (fact) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (valid[j]) {
                        const float tc = (cur_s[j] - smid) * scale;
                        sh.u[dst[j]] = ex2_approx(tc);
                        sh.v[dst[j]] = ex2_approx(-tc);
                        sh.g[dst[j]] = gain[j];
                    }
                }
                if (lane < 4) {                                  // zero tail: p = 0, w = 0 for the last block's padding
                    sh.u[m + lane] = 0.f;
                    sh.v[m + lane] = 0.f;
                    sh.g[m + lane] = 0.f;
                }
                __syncwarp();
                if (m > 96) val = wps_slate<4>(p, sh, m, lane);
                else if (m > 64) val = wps_slate<3>(p, sh, m, lane);
                else val = wps_slate<2>(p, sh, m, lane);          // (short slates leave lanes idle)
                __syncwarp();
                if (p.grad != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (32 * j + lane < n)
                            p.grad[(size_t)b * n + 32 * j + lane] = valid[j] ? sh.out[dst[j]] : 0.f;   // padded docs: exactly 0
                }
            } else {
                // scaled score range beyond the fp32 exponent budget of u*v (rare): direct ex2 form per ordered pair
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (valid[j]) sh.u[dst[j]] = cur_s[j];
                __syncwarp();
                float part = 0.f, cj[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cj[j] = 0.f;
                    if (valid[j]) {
                        const float pos1 = 2.0f + direct_rank_sum(sh.u, m, scale, cur_s[j]);
                        const float il2 = rcp_approx(lg2_approx(pos1));
                        part += gain[j] * il2;
                        cj[j] = (p.inv_B * (1.0f / kLn2)) * gain[j] * il2 * il2 * rcp_approx(pos1);
                        sh.g[dst[j]] = cj[j];
                    }
                }
                val = warp_sum(part);
                __syncwarp();
                if (p.grad != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (32 * j + lane < n) {
                            float acc = 0.f;
                            if (valid[j]) acc = direct_grad_sum(sh.u, sh.g, m, scale, cur_s[j], cj[j]);
                            p.grad[(size_t)b * n + 32 * j + lane] = p.alpha * acc;
                        }
                    }
                }
            }
            if (lane == 0) {
                p.slate_out[b] = val;
                wps_carry_add(cy, val);
            }
            __syncwarp();                                        // this slate's shared-memory reads are done
        }
        b = next_b;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cur_s[j] = next_s[j];
            cur_y[j] = next_y[j];
        }
    }

    // ---- batch loss: the CTA's four warps meet once, then ONE 64-bit atomic per CTA carries the fixed-point sum
    //      (low 52 bits) and counts the CTA as finished (bits 52..62); bit 63 is a sticky "a slate value was not
    //      finite" flag.  The CTA that completes the count holds the total in the value its own atomic returned ----
    if (lane == 0) {
        cta_fx[warp] = cy.fx;
        cta_bad[warp] = cy.nonfinite;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(p.ws + 2);
        long long fx = 0;
        int bad = 0;
#pragma unroll
        for (int w = 0; w < WPS_WARPS; ++w) fx += cta_fx[w], bad |= cta_bad[w];
        if (bad) atomicOr(acc64, 1ull << 63);                        // same address, same thread: ordered before the add
        const unsigned long long mine = (unsigned long long)fx + (1ull << kFixCountShift);
        const unsigned long long after = atomicAdd(acc64, mine) + mine;
        if (((after >> kFixCountShift) & 0x7ffull) == (unsigned long long)gridDim.x) {
            const double total = (double)(after & ((1ull << kFixCountShift) - 1ull));
            const double mean = total / ((double)kWpsFixScale * (double)p.B);
            p.loss_out[0] = (after >> 63) ? __int_as_float(0x7fc00000) : (float)(-mean);       // approxNDCG.py:53
            p.ws[0] = 0u;
            *acc64 = 0ull;
        }
    }
}

}  // namespace ltr
