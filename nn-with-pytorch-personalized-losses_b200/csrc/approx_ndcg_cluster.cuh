// K1 for long slates (128 < n <= 1024): one thread-block CLUSTER of 8 CTAs per slate
// (replaces losses/approxNDCG.py:7-53 of the reference; BASELINE configs[4]: 1024-doc slates).
//
// The pair work of a slate is the same symmetric schedule as the fast path -- every UNORDERED pair once per pass,
// sigma_ij = rcp(1 + u_i v_j) -- organised in groups of 128 compacted docs: group pair (g, g) is the block circulant
// of approx_ndcg_sym.cuh (ring of ceil(m_g / 4) lanes, steps 1 .. T/2 plus the in-block triangle), group pair
// (g1 < g2) a full ring of 32 lanes with rows from g1 and columns from g2.  All tile steps of a slate (16 G^2 + G
// of them for G groups) are dealt in equal contiguous ranges to the 32 warps of the cluster.  Every warp adds what it
// found for a doc (as a row or as a column) into its OWN [1024]-float partial array in shared memory; after a
// cluster barrier the CTA that owns a doc sums the 32 partials over distributed shared memory in a fixed order
// (bitwise reproducible), turns soft ranks into d loss / d pos, and the gradient pass repeats the walk.
// Each CTA prepares the whole slate itself (compaction, statistics, u / v / gains: 1024 docs are cheap next to
// 0.5 M pairs), so no broadcast is needed before the first pass.
#pragma once
#include <cooperative_groups.h>

#include "approx_ndcg_sym.cuh"

namespace ltr {
namespace cg = cooperative_groups;

constexpr int CL_CTAS = 8;
#ifndef CL_THREADS_OVERRIDE
#define CL_THREADS_OVERRIDE 128          // measured (64 / 512 slates of 1024 docs): 128 threads 43.6 / 239 us, 256 threads 46.4 / 303 us
#endif
constexpr int CL_THREADS = CL_THREADS_OVERRIDE;   // 4 warps: 32 warps share the tile steps of a slate
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_MAXN = 1024;
constexpr int CL_GROUP = 128;            // docs per group (32 lanes x 4)
constexpr int CL_PER_THREAD = CL_MAXN / CL_THREADS;   // positions a thread looks at while the slate is prepared

struct ClShared {
    float u[CL_MAXN];                    // 2^(t - c) of the compacted docs (direct form: raw score, -inf beyond m)
    float v[CL_MAXN];                    // 2^-(t - c)                      (direct form: raw score, +inf beyond m)
    float g[CL_MAXN];                    // normalised gains G_i
    float c[CL_MAXN];                    // d loss / d pos_i of every doc (own share computed here, the rest gathered)
    float part[CL_WARPS][CL_MAXN];       // per-warp partial sums (ranks, then gradients)
    short idx[CL_MAXN];                  // original position of the compacted doc
    float red[CL_WARPS];
    int ired[CL_WARPS][4];
    float value;                         // this CTA's share of the slate value
    int last;
};

// ---- pair tiles on 4 rows x 4 columns (FACT: approx_ndcg_sym.cuh's; direct form: p = 2^min(scale (s_i - s_j), 126)) ----
template <bool FACT>
__device__ __forceinline__ void cl_rank_tile(const float (&urow)[4], const float (&vcol)[4], float scale, float h,
                                             float (&acc_r)[4], float (&acc_r_hi)[4], float (&acc_c)[4]) {
    if (FACT) {
        rank_tile<4>(urow, vcol, h, acc_r, acc_r_hi, acc_c);
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float rr = rcp_approx(1.0f + ex2_capped((urow[r] - vcol[c]) * scale));
                acc_r[r] = fmaf(rr, h, acc_r[r]);
                acc_c[c] = fmaf(rr, h, acc_c[c]);
            }
    }
}
template <bool FACT>
__device__ __forceinline__ void cl_grad_tile(const float (&urow)[4], const float (&vcol)[4], const float (&crow)[4],
                                             const float (&ccol)[4], float scale, float (&acc_r)[4],
                                             float (&acc_r_hi)[4], float (&acc_c)[4]) {
    if (FACT) {
        grad_tile<4>(urow, vcol, crow, ccol, acc_r, acc_r_hi, acc_c);
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float pp = ex2_capped((urow[r] - vcol[c]) * scale);
                const float rr = rcp_approx(pp + 1.0f);
                const float e = (rr * (pp * rr)) * (ccol[c] - crow[r]);
                acc_r[r] += e;
                acc_c[c] -= e;
            }
    }
}

// One pass (GRAD = false: soft ranks, true: gradients) of this warp's range of tile steps.
//   steps of group pair (g, g):   D = 0 (in-block triangle), 1 .. T_g/2      -> T_g/2 + 1 steps
//   steps of group pair (g1, g2): D = 0 .. 31                                 -> 32 steps
template <bool FACT, bool GRAD>
__device__ __forceinline__ void cl_pass(ClShared& sh, float* mypart, int m, int groups, int first, int count, float scale) {
    const int lane = threadIdx.x & 31;
    int offset = 0;                                        // global index of the current pair's first step
    for (int g1 = 0; g1 < groups && count > 0; ++g1) {
        const int m1 = min(CL_GROUP, m - g1 * CL_GROUP);
        const int T1 = (m1 + 3) >> 2;                      // row blocks of g1
        for (int g2 = g1; g2 < groups && count > 0; ++g2) {
            const bool sym = g1 == g2;
            const int m2 = min(CL_GROUP, m - g2 * CL_GROUP);
            const int T2 = (m2 + 3) >> 2;
            const int ring = sym ? T1 : 32;
            const int S = ring >> 1;
            const int nsteps = sym ? S + 1 : 32;
            int lo = first - offset, hi = lo + count;      // this warp's steps inside the pair
            offset += nsteps;
            if (lo >= nsteps) continue;
            if (lo < 0) lo = 0;                            // (cannot happen: ranges are consumed in order)
            if (hi > nsteps) hi = nsteps;
            const int done = hi - lo;
            first += done;
            count -= done;

            const bool active = lane < ring;
            const int src = active ? ((lane + 1 == ring) ? 0 : lane + 1) : lane;
            const int rbase = g1 * CL_GROUP + lane * 4, cgrp = g2 * CL_GROUP;
            float urow[4], vrow[4], crow[4], vcol[4], ccol[4], acc_r[4], acc_r_hi[4], acc_c[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                urow[r] = sh.u[rbase + r];                 // (beyond m: zero / -inf tail, see the preparation)
                vrow[r] = sh.v[rbase + r];
                crow[r] = GRAD ? sh.c[rbase + r] : 0.f;
                acc_r[r] = 0.f;
                acc_r_hi[r] = 0.f;
                acc_c[r] = 0.f;
            }
            int D = lo;
            if (sym && D == 0) {                           // pairs inside the lane's own block
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = r + 1; c < 4; ++c) {
                        const float pp = FACT ? urow[r] * vrow[c] : ex2_capped((urow[r] - vrow[c]) * scale);
                        const float rr = rcp_approx(pp + 1.0f);
                        if (!GRAD) {
                            acc_r[r] += rr;
                            acc_r[c] = fmaf(pp, rr, acc_r[c]);
                        } else {
                            const float e = (rr * (pp * rr)) * (crow[c] - crow[r]);
                            acc_r[r] += e;
                            acc_r[c] -= e;
                        }
                    }
                D = 1;
            }
            int last_col = -1;                             // column block the accumulators belong to after the loop
            if (D < hi) {
                int cb = lane + D;
                if (cb >= ring) cb -= ring;
                if (!active) cb = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    vcol[r] = sh.v[cgrp + cb * 4 + r];
                    ccol[r] = GRAD ? sh.c[cgrp + cb * 4 + r] : 0.f;
                }
#pragma unroll 1
                for (; D < hi; ++D) {
                    const bool half = sym && ((ring & 1) == 0) && D == S;      // antipodal step: pairs met twice
                    if (!GRAD) {
                        cl_rank_tile<FACT>(urow, vcol, scale, half ? 0.5f : 1.0f, acc_r, acc_r_hi, acc_c);
                    } else if (half) {
                        float ch[4], cch[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) ch[r] = 0.5f * crow[r], cch[r] = 0.5f * ccol[r];
                        cl_grad_tile<FACT>(urow, vcol, ch, cch, scale, acc_r, acc_r_hi, acc_c);
                    } else {
                        cl_grad_tile<FACT>(urow, vcol, crow, ccol, scale, acc_r, acc_r_hi, acc_c);
                    }
                    if (D + 1 < hi) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                            if (GRAD) ccol[r] = __shfl_sync(LTR_FULL_MASK, ccol[r], src);
                            acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
                        }
                    }
                }
                last_col = lane + hi - 1;
                if (last_col >= ring) last_col -= ring;
            }
            // partial sums into this warp's own array (same warp, different lanes at different times: ordered below)
            if (lane < T1) {
#pragma unroll
                for (int r = 0; r < 4; ++r) mypart[rbase + r] += acc_r[r] + acc_r_hi[r];
            }
            __syncwarp();
            if (active && last_col >= 0 && last_col < T2) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float* slot = mypart + cgrp + last_col * 4 + r;
                    *slot = GRAD ? *slot + acc_c[r] : *slot - acc_c[r];     // ranks: sigma_ji = 1 - sigma_ij, see cl_finish_rank
                }
            }
            __syncwarp();
        }
    }
}

template <bool UNUSED = true>
__global__ void __launch_bounds__(CL_THREADS, 2) approx_ndcg_cluster_kernel(const ApproxParams p) {
    extern __shared__ __align__(16) unsigned char cl_smem[];
    ClShared& sh = *reinterpret_cast<ClShared*>(cl_smem);
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();            // 0 .. 7
    const int b = (int)blockIdx.x / CL_CTAS;               // slate of this cluster
    const int n = p.n;
    const float scale = p.alpha * kLog2e;
    const float inf = __int_as_float(0x7f800000);
    const float* srow = p.s + (size_t)b * n;
    const float* yrow = p.y + (size_t)b * n;

    // ---------------- preparation (every CTA, whole slate): compaction, statistics, gains, u / v ----------------
    // thread t looks at positions CL_PER_THREAD t .. CL_PER_THREAD (t + 1) - 1
    float sv[CL_PER_THREAD], yv[CL_PER_THREAD];
    int cnt = 0, kmin = 0x7fffffff, kmax = (int)0x80000000;
    unsigned lvl_lo = 0, lvl_hi = 0;                       // counts of labels 1, 2 (16 bits each) and 3, 4
    bool odd = false;
#pragma unroll
    for (int e = 0; e < CL_PER_THREAD; ++e) {
        const int i = CL_PER_THREAD * tid + e;
        sv[e] = 0.f;
        yv[e] = p.pad;
        if (i < n) {
            sv[e] = ld_stream(srow + i);
            yv[e] = ld_stream(yrow + i);
        }
        const bool valid = (i < n) && (yv[e] != p.pad);
        if (valid) {
            ++cnt;
            const int key = float_order_key(sv[e]);
            kmin = min(kmin, key);
            kmax = max(kmax, key);
            const float yc = fmaxf(yv[e], 0.f);                                   // approxNDCG.py:37-38 clamp
            const int lv = (int)fminf(yc, 100.f);
            if (yc == (float)lv && lv <= SYM_MAX_LABEL) {
                lvl_lo += (lv == 1) ? 1u : ((lv == 2) ? 0x10000u : 0u);
                lvl_hi += (lv == 3) ? 1u : ((lv == 4) ? 0x10000u : 0u);
            } else {
                odd = true;
            }
        } else if (i < n && rank == 0 && p.grad != nullptr) {
            p.grad[(size_t)b * n + i] = 0.f;                                      // padded docs: exactly 0
        }
    }
    // exclusive prefix of cnt over the 128 threads + block totals
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(LTR_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    kmin = __reduce_min_sync(LTR_FULL_MASK, kmin);
    kmax = __reduce_max_sync(LTR_FULL_MASK, kmax);
    lvl_lo = __reduce_add_sync(LTR_FULL_MASK, lvl_lo);
    lvl_hi = __reduce_add_sync(LTR_FULL_MASK, lvl_hi);
    odd = __any_sync(LTR_FULL_MASK, odd);
    if (lane == 31) {
        sh.ired[warp][0] = incl;
        sh.ired[warp][1] = kmin;
        sh.ired[warp][2] = kmax;
        sh.ired[warp][3] = odd ? 1 : 0;
        sh.red[warp] = __uint_as_float(lvl_lo);
        sh.c[warp] = __uint_as_float(lvl_hi);              // (scratch: sh.c is written later)
    }
    __syncthreads();
    int m = 0, off = 0, any_odd = 0;
    unsigned lo_cnt = 0, hi_cnt = 0;
    kmin = 0x7fffffff;
    kmax = (int)0x80000000;
#pragma unroll
    for (int w = 0; w < CL_WARPS; ++w) {
        off += (w < warp) ? sh.ired[w][0] : 0;
        m += sh.ired[w][0];
        kmin = min(kmin, sh.ired[w][1]);
        kmax = max(kmax, sh.ired[w][2]);
        any_odd |= sh.ired[w][3];
        lo_cnt += __float_as_uint(sh.red[w]);
        hi_cnt += __float_as_uint(sh.c[w]);
    }
    __syncthreads();
    int dst = off + incl - cnt;                            // compacted position of this thread's first valid doc

    if (m == 0) {                                          // fully padded slate: value 0, no gradient
        if (rank == 0 && tid == 0) p.slate_out[b] = 0.f;
    }
    const int groups = (m + CL_GROUP - 1) / CL_GROUP;

    // ideal DCG over ALL positions (approxNDCG.py:41-43)
    float maxdcg;
    if (!any_odd) {
        const int e4 = (int)(hi_cnt >> 16), e3 = e4 + (int)(hi_cnt & 0xffffu), e2 = e3 + (int)(lo_cnt >> 16),
                  e1 = e2 + (int)(lo_cnt & 0xffffu);
        const double dcg = 15.0 * p.disc_prefix[e4] + 7.0 * (p.disc_prefix[e3] - p.disc_prefix[e4]) +
                           3.0 * (p.disc_prefix[e2] - p.disc_prefix[e3]) + (p.disc_prefix[e1] - p.disc_prefix[e2]);
        maxdcg = fmaxf((float)dcg, p.eps);
    } else {
        // general labels: every doc finds its ideal rank by counting (rare path); labels staged in sh.c
        int d = dst;
#pragma unroll
        for (int e = 0; e < CL_PER_THREAD; ++e)
            if (CL_PER_THREAD * tid + e < n && yv[e] != p.pad) sh.c[d++] = fmaxf(yv[e], 0.f);
        __syncthreads();
        float part = 0.f;
        d = dst;
#pragma unroll 1
        for (int e = 0; e < CL_PER_THREAD; ++e) {
            if (CL_PER_THREAD * tid + e < n && yv[e] != p.pad) {
                const float y = fmaxf(yv[e], 0.f);
                int rk = 0;
                for (int k = 0; k < m; ++k) {
                    const float o = sh.c[k];
                    rk += (o > y || (o == y && k < d)) ? 1 : 0;
                }
                part += gain_exp2(y) / log2f(2.0f + (float)rk);
                ++d;
            }
        }
        part = warp_sum(part);
        __syncthreads();
        if (lane == 0) sh.red[warp] = part;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < CL_WARPS; ++w) tot += sh.red[w];
        maxdcg = fmaxf(tot, p.eps);
        __syncthreads();
    }
    const float smin = float_from_key(kmin), smax = float_from_key(kmax);
    const float range = (smax - smin) * scale;
    const bool fact = fabsf(range) < 126.0f;               // false for inf / NaN ranges too
    const float smid = 0.5f * smin + 0.5f * smax;
    {
        int d = dst;
#pragma unroll
        for (int e = 0; e < CL_PER_THREAD; ++e) {
            const int i = CL_PER_THREAD * tid + e;
            if (i < n && yv[e] != p.pad) {
                const float yc = fmaxf(yv[e], 0.f);
                const float tc = (sv[e] - smid) * scale;
                sh.u[d] = fact ? ex2_approx(tc) : sv[e];
                sh.v[d] = fact ? ex2_approx(-tc) : sv[e];
                sh.g[d] = __fdividef(gain_exp2(yc), maxdcg);                      // G_i (approxNDCG.py:44)
                sh.idx[d] = (short)i;
                ++d;
            }
        }
        // tail up to the next multiple of 128: p = 0 (sigma_ij = 1, w = 0) for every pair with a doc that does not exist
        for (int i = m + tid; i < groups * CL_GROUP; i += CL_THREADS) {
            sh.u[i] = fact ? 0.f : -inf;
            sh.v[i] = fact ? 0.f : inf;
            sh.g[i] = 0.f;
            sh.c[i] = 0.f;
        }
        for (int i = lane; i < CL_MAXN; i += 32) sh.part[warp][i] = 0.f;          // (each warp clears its own array)
    }
    __syncthreads();

    // ---------------- rank pass: this warp's share of the slate's tile steps ----------------
    int total = 0;
    for (int g = 0; g < groups; ++g) {
        const int mg = min(CL_GROUP, m - g * CL_GROUP);
        total += (((mg + 3) >> 2) >> 1) + 1 + 32 * (groups - 1 - g);
    }
    const int gw = rank * CL_WARPS + warp;                 // warp index inside the cluster
    const int first = (int)(((long long)total * gw) / (CL_CTAS * CL_WARPS));
    const int count = (int)(((long long)total * (gw + 1)) / (CL_CTAS * CL_WARPS)) - first;
    float* mypart = sh.part[warp];
    if (fact) cl_pass<true, false>(sh, mypart, m, groups, first, count, scale);
    else cl_pass<false, false>(sh, mypart, m, groups, first, count, scale);
    cluster.sync();

    // ---------------- the owner of a doc sums its 32 partials, soft rank -> d loss / d pos, slate value ----------------
    const int last_m = m - (groups - 1) * CL_GROUP;        // docs of the last group
    const int last_T = (last_m + 3) >> 2;
    float val_part = 0.f;
    {
        const int d = rank * CL_GROUP + tid;               // this thread's doc (threads 0 .. 127)
        float ci = 0.f;
        if (tid < CL_GROUP && d < m) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < CL_CTAS; ++r) {
                const ClShared* other = cluster.map_shared_rank(&sh, r);
#pragma unroll
                for (int w = 0; w < CL_WARPS; ++w) sum += other->part[w][d];
            }
            // column side: sum sigma_ji = (#rows met) - sum sigma_ij; rows met: 128 per earlier group + 4 (T - 1) / 2
            // inside the own group.  Row side: every doc that does not exist reads sigma_ij = 1.
            const int g = rank;                            // (group == owner rank by construction)
            const int Tg = (g == groups - 1) ? last_T : 32;
            float fix = 0.f;
            if (g == groups - 1) {
                const int npadc = Tg * 4 - last_m;
                const int e = Tg - 1 - (tid >> 2);             // (tid < 128: block of the doc inside its group)
                const float f = (2 * e == Tg) ? 0.5f : ((2 * e < Tg) ? 1.0f : 0.0f);
                fix = f * (float)npadc;
            } else {
                fix = (float)(CL_GROUP - last_m);          // the tail of the last group, met once in the (g, last) ring
            }
            sum = (sum - fix) + (float)(CL_GROUP * g) + 0.5f * (float)(4 * (Tg - 1));
            const float pos1 = 2.0f + sum;                 // 1 + pos_i
            const float il2 = rcp_approx(lg2_approx(pos1));
            const float gi = sh.g[d] * il2;
            val_part = gi;
            ci = (p.inv_B * (1.0f / kLn2)) * gi * il2 * rcp_approx(pos1);
        }
        val_part = warp_sum(val_part);
        if (lane == 0) sh.red[warp] = val_part;
        __syncthreads();
        if (tid == 0) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < CL_WARPS; ++w) v += sh.red[w];
            sh.value = v;
        }
        // (every remote read of the rank partials must be over before anybody clears them: barrier below)
        cluster.sync();
        if (tid < CL_GROUP && rank * CL_GROUP + tid < groups * CL_GROUP) sh.c[rank * CL_GROUP + tid] = ci;
        for (int i = lane; i < CL_MAXN; i += 32) sh.part[warp][i] = 0.f;
    }
    cluster.sync();
    if (rank == 0 && tid == 0 && m > 0) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < CL_CTAS; ++r) v += cluster.map_shared_rank(&sh, r)->value;
        p.slate_out[b] = v;
    }
    // gather every doc's d loss / d pos
    for (int d = tid; d < groups * CL_GROUP; d += CL_THREADS) {
        const int owner = d / CL_GROUP;
        if (owner != rank) sh.c[d] = cluster.map_shared_rank(&sh, owner)->c[d];
    }
    __syncthreads();

    // ---------------- gradient pass ----------------
    if (p.grad != nullptr && m > 0) {
        if (fact) cl_pass<true, true>(sh, mypart, m, groups, first, count, scale);
        else cl_pass<false, true>(sh, mypart, m, groups, first, count, scale);
    }
    cluster.sync();
    if (p.grad != nullptr) {
        const int d = rank * CL_GROUP + tid;
        if (tid < CL_GROUP && d < m) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < CL_CTAS; ++r) {
                const ClShared* other = cluster.map_shared_rank(&sh, r);
#pragma unroll
                for (int w = 0; w < CL_WARPS; ++w) sum += other->part[w][d];
            }
            p.grad[(size_t)b * n + sh.idx[d]] = p.alpha * sum;
        }
    }
    cluster.sync();                                        // nobody exits while its shared memory may still be read

    // ---- batch loss: the last cluster to finish adds the slate values in index order ----
    if (rank == 0) {
        if (tid == 0) {
            __threadfence();
            const unsigned done = atomicAdd(&p.ws[1], 1u);
            sh.last = (done == (unsigned)p.B - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (sh.last) {
            __threadfence();
            double acc = 0.0;
            for (int i = tid; i < p.B; i += CL_THREADS) acc += (double)__ldcg(p.slate_out + i);
            acc = warp_sum(acc);
            __shared__ double dred[CL_WARPS];
            if (lane == 0) dred[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < CL_WARPS; ++w) t += dred[w];
                p.loss_out[0] = (float)(-t / (double)p.B);       // approxNDCG.py:53
                p.ws[1] = 0u;
                __threadfence();
            }
        }
    }
}

}  // namespace ltr
