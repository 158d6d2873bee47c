// K1: fused forward+backward approxNDCG (replaces losses/approxNDCG.py:7-53 of the reference).
//
// One CTA works on one slate at a time (persistent CTAs pull slates from an atomic counter).
// Valid (non-padded) docs are compacted to the front of shared memory -- the loss is
// permutation invariant, so no score sort is needed.  The O(m^2) pairwise
// sigmoid sigma_ij = 1 / (1 + 2^(t_i - t_j)), t = alpha*log2(e)*s, is evaluated in registers:
//   * factorised form  sigma_ij = rcp(1 + u_i * v_j),  u_i = 2^(t_i - c), v_j = 2^-(t_j - c)
//     -> one FFMA + one MUFU.RCP per ordered pair (used when the slate's score range keeps
//        u, v inside fp32 range; otherwise the direct ex2 form is used);
//   * pass 1 accumulates soft ranks pos_i, pass 2 the analytic gradient
//        d loss/d s_k = alpha * sum_j w_kj (c_j - c_k),  w = sigma(1-sigma),
//        c_i = G_i / (B ln2 (1+pos_i) log2(1+pos_i)^2)          (SURVEY A.1).
// The B x n x n tensors of the reference never exist.
#include "approx_ndcg_sym.cuh"
#include "approx_ndcg_warp.cuh"
#include "approx_ndcg_cluster.cuh"
#include "ltr_common.cuh"
#include <stdlib.h>

#include "ltr_host.h"

namespace ltr {

// p = 2^(t_i - t_j) for one (row, column) in either formulation; sigma_ij = 1/(1+p), sigma_ji = p/(1+p)
template <bool FACT>
__device__ __forceinline__ float pair_p(float row, float col, float scale) {
    if (FACT) return row * col;                                       // row = u_i, col = v_j
    return ex2_capped((row - col) * scale);                           // row = s_i, col = s_j (raw scores)
}

// PHASE 0: both passes over all rows of the slate (one CTA per slate).  PHASE 1 / 2: the rank pass / the gradient
// pass over the rows [row_lo, row_hi) only -- long slates are split over several CTAs (and two launches: the
// gradient pass needs every doc's d loss / d pos, exchanged through p.cbuf).
template <int THREADS, bool FACT, bool CLAMP, int PHASE>
__device__ __forceinline__ void pair_passes(const ApproxParams& p, int b, int m, float scale, const float* sh_row,
                                            const float* sh_col, float* sh_c, const float* sh_g,
                                            const int* sh_idx, float* red, int row_lo, int row_hi) {
    const int tid = threadIdx.x;
    const int m4 = (m + 3) & ~3;
    const float eps = p.eps;
    if (row_hi > m) row_hi = m;
    // ---------------- pass 1: soft ranks, per-doc dloss/dpos, slate value -----------------
    float ndcg_part = 0.f;
    if (PHASE != 2)
    for (int i = row_lo + tid; i < row_hi; i += THREADS) {
        const float r = sh_row[i];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int j = 0; j < m4; j += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(sh_col + j);
            float s0 = rcp_approx(1.0f + pair_p<FACT>(r, c4.x, scale)), s1 = rcp_approx(1.0f + pair_p<FACT>(r, c4.y, scale));
            float s2 = rcp_approx(1.0f + pair_p<FACT>(r, c4.z, scale)), s3 = rcp_approx(1.0f + pair_p<FACT>(r, c4.w, scale));
            if (CLAMP) {
                s0 = fmaxf(s0, eps); s1 = fmaxf(s1, eps); s2 = fmaxf(s2, eps); s3 = fmaxf(s3, eps);
            }
            a0 += s0; a1 += s1; a2 += s2; a3 += s3;
        }
        float self = rcp_approx(1.0f + pair_p<FACT>(r, sh_col[i], scale));
        if (CLAMP) self = fmaxf(self, eps);
        float acc = (a0 + a1) + (a2 + a3) - self;
        // padding columns: the factorised form reads sigma = 1 there (v = 0), the direct form 0 (clamped up to eps)
        if (FACT) acc -= (float)(m4 - m);
        else if (CLAMP) acc -= (float)(m4 - m) * eps;
        const float pos = 1.0f + acc;
        const float l2 = log2f(1.0f + pos);
        const float g = sh_g[i];
        ndcg_part += g / l2;
        const float ci = p.inv_B * g / (kLn2 * (1.0f + pos) * l2 * l2);
        if (PHASE == 1) p.cbuf[(size_t)b * p.n + i] = ci;
        else sh_c[i] = ci;
    }
    if (PHASE != 2) {
        const float ndcg = block_sum<THREADS>(ndcg_part, red);   // also orders sh_c writes before pass 2
        if (tid == 0) {
            if (PHASE == 1) p.part[(size_t)b * gridDim.x + blockIdx.x] = ndcg;
            else p.slate_out[b] = ndcg;
        }
    }
    if (PHASE == 1 || p.grad == nullptr) return;
    __syncthreads();
    // ---------------- pass 2: analytic gradient ---------------------------------------------
    for (int i = row_lo + tid; i < row_hi; i += THREADS) {
        const float r = sh_row[i];
        float A0 = 0.f, A1 = 0.f, W0 = 0.f, W1 = 0.f;
#pragma unroll 4
        for (int j = 0; j < m4; j += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(sh_col + j);
            const float4 k4 = *reinterpret_cast<const float4*>(sh_c + j);
            const float p0 = pair_p<FACT>(r, c4.x, scale), p1 = pair_p<FACT>(r, c4.y, scale);
            const float p2 = pair_p<FACT>(r, c4.z, scale), p3 = pair_p<FACT>(r, c4.w, scale);
            const float s0 = rcp_approx(1.0f + p0), s1 = rcp_approx(1.0f + p1);
            const float s2 = rcp_approx(1.0f + p2), s3 = rcp_approx(1.0f + p3);
            const float q0 = p0 * s0, q1 = p1 * s1, q2 = p2 * s2, q3 = p3 * s3;      // sigma_ji
            // w = sigma_ij * sigma_ji: the sigmoid derivative without the 1 - sigma cancellation
            const float w0 = s0 * q0, w1 = s1 * q1, w2 = s2 * q2, w3 = s3 * q3;
            if (CLAMP) {
                // row term: d pos_i / d s_i passes only where sigma_ij >= eps;
                // column term: d pos_j / d s_i passes only where sigma_ji >= eps
                A0 = fmaf((q0 >= eps) ? w0 : 0.f, k4.x, A0);
                A1 = fmaf((q1 >= eps) ? w1 : 0.f, k4.y, A1);
                A0 = fmaf((q2 >= eps) ? w2 : 0.f, k4.z, A0);
                A1 = fmaf((q3 >= eps) ? w3 : 0.f, k4.w, A1);
                W0 += (s0 >= eps) ? w0 : 0.f;
                W1 += (s1 >= eps) ? w1 : 0.f;
                W0 += (s2 >= eps) ? w2 : 0.f;
                W1 += (s3 >= eps) ? w3 : 0.f;
            } else {
                A0 = fmaf(w0, k4.x, A0); A1 = fmaf(w1, k4.y, A1);
                A0 = fmaf(w2, k4.z, A0); A1 = fmaf(w3, k4.w, A1);
                W0 += w0; W1 += w1; W0 += w2; W1 += w3;
            }
        }
        // the j == i term is w_ii (c_i - c_i) = 0 and needs no special case (without the clamp asymmetry)
        const float ci = sh_c[i];
        float gsum = (A0 + A1) - ci * (W0 + W1);
        // (with the clamp, the self pair sigma_ii = 0.5 passes both masks and still cancels)
        p.grad[(size_t)b * p.n + sh_idx[i]] = p.alpha * gsum;
    }
}

template <int THREADS, int PHASE>
__global__ void __launch_bounds__(THREADS) approx_ndcg_kernel(const ApproxParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int W = THREADS / 32;
    const int npad = p.npad;
    float* sh_a = smem;              // scaled scores t (direct form) or u (factorised form)
    float* sh_v = sh_a + npad;       // v (factorised form) / sort scratch
    float* sh_c = sh_v + npad;       // dloss/dpos per doc       / sort scratch (cont.)
    float* sh_g = sh_c + npad;       // labels, then normalised gains G
    int* sh_idx = reinterpret_cast<int*>(sh_g + npad);   // original position of compacted doc
    __shared__ float red[4 * 32];
    __shared__ int sh_wcnt[32];
    __shared__ int sh_lvl[16];
    __shared__ int sh_next;
    __shared__ float sh_maxdcg;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const float scale = p.alpha * kLog2e;
    const float inf = __int_as_float(0x7f800000);

    for (;;) {
        if (PHASE == 0 && tid == 0) sh_next = (int)atomicAdd(&p.ws[0], 1u);
        if (tid < 16) sh_lvl[tid] = 0;
        __syncthreads();
        const int b = (PHASE == 0) ? sh_next : (int)blockIdx.y;
        if (b >= p.B) break;
        const float* srow = p.s + (size_t)b * n;
        const float* yrow = p.y + (size_t)b * n;

        // ---- load, compact valid docs to the front, gather per-slate statistics --------------
        int m = 0;
        float tmin = inf, tmax = -inf, lmax = 0.f;
        int all_int = 1;
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            float sv = 0.f, yv = p.pad;
            if (i < n) {
                sv = ld_stream(srow + i);
                yv = ld_stream(yrow + i);
            }
            const bool valid = (i < n) && (yv != p.pad);
            const unsigned bal = __ballot_sync(LTR_FULL_MASK, valid);
            if (lane == 0) sh_wcnt[warp] = __popc(bal);
            __syncthreads();
            int off = m, tot = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int c = sh_wcnt[w];
                off += (w < warp) ? c : 0;
                tot += c;
            }
            if (valid) {
                const int pos = off + __popc(bal & ((1u << lane) - 1u));
                const float t = sv;                         // raw score; scaled per pair / per doc below
                const float yc = fmaxf(yv, 0.f);            // approxNDCG.py:37-38 clamp
                sh_a[pos] = t;
                sh_g[pos] = yc;
                sh_idx[pos] = i;
                tmin = fminf(tmin, t);
                tmax = fmaxf(tmax, t);
                lmax = fmaxf(lmax, yc);
                all_int &= (yc == floorf(yc)) ? 1 : 0;
            } else if (i < n && p.grad != nullptr && PHASE != 1 && (PHASE == 0 || blockIdx.x == 0)) {
                p.grad[(size_t)b * n + i] = 0.f;            // padded docs get exactly 0
            }
            m += tot;
            __syncthreads();
        }
        if (m == 0) {                                        // fully padded slate: value 0, no gradient
            if (tid == 0) {
                if (PHASE == 0) p.slate_out[b] = 0.f;
                if (PHASE == 1) p.part[(size_t)b * gridDim.x + blockIdx.x] = 0.f;
            }
            if (PHASE == 0) continue;
            return;
        }
        // block reduce (min, max, label max, all-int flag)
        tmin = warp_min(tmin);
        tmax = warp_max(tmax);
        lmax = warp_max(lmax);
        all_int = __all_sync(LTR_FULL_MASK, all_int);
        if (lane == 0) {
            red[warp] = tmin;
            red[32 + warp] = tmax;
            red[64 + warp] = lmax;
            red[96 + warp] = (float)all_int;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < W; ++w) {
            tmin = fminf(tmin, red[w]);
            tmax = fmaxf(tmax, red[32 + w]);
            lmax = fmaxf(lmax, red[64 + w]);
            all_int &= (int)red[96 + w];
        }
        const int m4 = (m + 3) & ~3;

        // ---- ideal DCG over ALL positions (approxNDCG.py:41-43) --------------------------------
        const bool hist_ok = all_int && lmax <= 15.f && p.disc_prefix != nullptr && m < p.disc_len;
        if (hist_ok) {
            // integer labels: count docs per level, each level occupies a contiguous rank range
            const int lmaxi = (int)lmax;
            for (int base = 0; base < m; base += THREADS) {
                const int i = base + tid;
                const int yi = (i < m) ? (int)sh_g[i] : -1;
                for (int L = 1; L <= lmaxi; ++L) {
                    const int c = __popc(__ballot_sync(LTR_FULL_MASK, yi == L));
                    if (lane == 0 && c) atomicAdd(&sh_lvl[L], c);
                }
            }
            __syncthreads();
            if (warp == 0) {
                double contrib = 0.0;
                if (lane >= 1 && lane <= lmaxi) {
                    int start = 0;
                    for (int L = lane + 1; L <= lmaxi; ++L) start += sh_lvl[L];
                    const int cnt = sh_lvl[lane];
                    contrib = (double)((float)(1u << lane) - 1.0f) * (p.disc_prefix[start + cnt] - p.disc_prefix[start]);
                }
                contrib = warp_sum(contrib);
                if (lane == 0) sh_maxdcg = fmaxf((float)contrib, p.eps);
            }
        } else {
            // general labels: bitonic sort (descending) in scratch aliased over sh_v/sh_c
            float* sortbuf = sh_v;
            const int len = next_pow2(m);
            for (int i = tid; i < len; i += THREADS) sortbuf[i] = (i < m) ? sh_g[i] : 0.f;
            block_bitonic_desc(sortbuf, len);
            float part = 0.f;
            for (int i = tid; i < m; i += THREADS) part += gain_exp2(sortbuf[i]) / log2f(2.0f + (float)i);
            const float tot = block_sum<THREADS>(part, red);
            if (tid == 0) sh_maxdcg = fmaxf(tot, p.eps);
        }
        __syncthreads();
        const float maxdcg = sh_maxdcg;

        // ---- per-doc quantities -----------------------------------------------------------------
        const float range = (tmax - tmin) * scale;
        const bool fact = (fabsf(range) < 126.0f);           // false for inf/NaN ranges too; keeps u*v finite
        const float center = 0.5f * tmax + 0.5f * tmin;
        const float sentinel = (scale >= 0.f) ? -inf : inf;  // direct form: 2^(scale*(x - pad)) saturates at 2^126
        for (int i = tid; i < m4 + 4 && i < npad; i += THREADS) {
            if (i < m) {
                sh_g[i] = gain_exp2(sh_g[i]) / maxdcg;       // G_i (approxNDCG.py:44)
                if (fact) {
                    const float tc = (sh_a[i] - center) * scale;
                    sh_a[i] = ex2_approx(tc);
                    sh_v[i] = ex2_approx(-tc);
                }
            } else {
                // padding columns: factorised p = 0 (sigma = 1, w = 0; counted out in pair_passes),
                // direct p = 2^126 (sigma ~ 0, w ~ 0)
                sh_a[i] = fact ? 0.0f : sentinel;
                sh_v[i] = 0.0f;
                sh_c[i] = 0.f;
                sh_g[i] = 0.f;
            }
        }
        __syncthreads();

        int row_lo = 0, row_hi = m;
        if (PHASE != 0) {
            row_lo = (int)blockIdx.x * p.rows_per_cta;
            row_hi = row_lo + p.rows_per_cta;
            if (PHASE == 2) {                                // every doc's d loss / d pos from the rank launch
                for (int i = tid; i < m; i += THREADS) sh_c[i] = p.cbuf[(size_t)b * n + i];
                __syncthreads();
            }
        }
        if (fact) {
            if (p.clamp) pair_passes<THREADS, true, true, PHASE>(p, b, m, scale, sh_a, sh_v, sh_c, sh_g, sh_idx, red, row_lo, row_hi);
            else         pair_passes<THREADS, true, false, PHASE>(p, b, m, scale, sh_a, sh_v, sh_c, sh_g, sh_idx, red, row_lo, row_hi);
        } else {
            if (p.clamp) pair_passes<THREADS, false, true, PHASE>(p, b, m, scale, sh_a, sh_a, sh_c, sh_g, sh_idx, red, row_lo, row_hi);
            else         pair_passes<THREADS, false, false, PHASE>(p, b, m, scale, sh_a, sh_a, sh_c, sh_g, sh_idx, red, row_lo, row_hi);
        }
        if (PHASE != 0) return;
        __syncthreads();
    }
    if (PHASE != 0) return;

    // ---- last CTA to finish reduces the per-slate values in a fixed order ------------------------
    __shared__ double dred[32];
    __syncthreads();   // every thread has read its last sh_next before it is reused as the "last CTA" flag
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&p.ws[1], 1u);
        sh_next = (done == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sh_next) {
        __threadfence();
        double acc = 0.0;
        for (int i = tid; i < p.B; i += THREADS) acc += (double)__ldcg(p.slate_out + i);
        acc = warp_sum(acc);
        if (lane == 0) dred[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < W; ++w) t += dred[w];
            p.loss_out[0] = (float)(-t / (double)p.B);       // approxNDCG.py:53
            p.ws[0] = 0u;
            p.ws[1] = 0u;
            __threadfence();
        }
    }
}

template <int THREADS>
static int launch_approx(const ApproxParams& p, cudaStream_t st) {
    const size_t smem = (size_t)p.npad * 5 * sizeof(float);
    static int occ_cached = 0;               // (occupancy is a property of the kernel image: the same on every B200)
    static size_t occ_smem = (size_t)-1;
    if (smem > 48 * 1024)                    // a per-device function attribute: set per call, not cached per process
        LTR_CUDA(cudaFuncSetAttribute(approx_ndcg_kernel<THREADS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    if (occ_smem != smem) {
        LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cached, approx_ndcg_kernel<THREADS, 0>, THREADS,
                                                               smem));
        occ_smem = smem;
    }
    int grid = device_sm_count() * (occ_cached > 0 ? occ_cached : 1);
    if (grid > p.B) grid = p.B;
    approx_ndcg_kernel<THREADS, 0><<<grid, THREADS, smem, st>>>(p);
    return (int)cudaGetLastError();
}

// slate value = sum of its chunks' partial values; loss = -mean (fixed reduction shape, one CTA)
__global__ void __launch_bounds__(256) approx_finish_kernel(const float* part, int nchunks, int B, float* slate_out,
                                                            float* loss_out) {
    __shared__ double red[8];
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += 256) {
        float v = 0.f;
        for (int c = 0; c < nchunks; ++c) v += part[(size_t)b * nchunks + c];
        slate_out[b] = v;
        acc += (double)v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        loss_out[0] = (float)(-t / (double)B);               // approxNDCG.py:53
    }
}

// Long slates, few of them: rows of a slate are split over ceil(n / 128) CTAs, rank launch then gradient launch.
static int launch_approx_split(ApproxParams p, cudaStream_t st) {
    constexpr int THREADS = 128;
    const size_t smem = (size_t)p.npad * 5 * sizeof(float);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        LTR_CUDA(cudaFuncSetAttribute(approx_ndcg_kernel<THREADS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LTR_CUDA(cudaFuncSetAttribute(approx_ndcg_kernel<THREADS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    p.rows_per_cta = THREADS;
    const int nchunks = (p.n + THREADS - 1) / THREADS;
    const dim3 grid(nchunks, p.B);
    approx_ndcg_kernel<THREADS, 1><<<grid, THREADS, smem, st>>>(p);
    if (p.grad) approx_ndcg_kernel<THREADS, 2><<<grid, THREADS, smem, st>>>(p);
    approx_finish_kernel<<<1, 256, 0, st>>>(p.part, nchunks, p.B, p.slate_out, p.loss_out);
    return (int)cudaGetLastError();
}

#ifndef WPS_MIN_BATCH
#define WPS_MIN_BATCH 16384    // batches of at least this many slates take the one-warp-per-slate kernel (coarser work
                               // items, fewer instructions per slate); smaller ones the four-warps-per-slate kernel
#endif

// persistent launch of a fast-path kernel: as many CTAs as are resident (or fewer for small batches), with
// programmatic stream serialization so that consecutive launches overlap ramp-up and tail
template <typename Kernel>
static int launch_persistent(Kernel kernel, int threads, int slates_per_cta, int per_sm_cap, const ApproxParams& p,
                             cudaStream_t st, int& occ_cached) {
    if (occ_cached == 0) {
        LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cached, kernel, threads, 0));
        if (occ_cached <= 0) occ_cached = 1;
    }
    int grid = device_sm_count() * ((per_sm_cap > 0 && per_sm_cap < occ_cached) ? per_sm_cap : occ_cached);
    const int need = (p.B + slates_per_cta - 1) / slates_per_cta;
    if (grid > need) grid = need;
    if (grid > 2047) grid = 2047;                                      // the finished-CTA count has 11 bits
#if SYM_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, p);
#else
    kernel<<<grid, threads, 0, st>>>(p);
    return (int)cudaGetLastError();
#endif
}

// long slates: one cluster of CL_CTAS CTAs per slate (approx_ndcg_cluster.cuh)
static int launch_approx_cluster(const ApproxParams& p, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.B * CL_CTAS);
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = sizeof(ClShared);
    cfg.stream = st;
    // a per-device function attribute: set on every call (cheap), never cached in a process-wide flag
    LTR_CUDA(cudaFuncSetAttribute(approx_ndcg_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(ClShared)));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL_CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, approx_ndcg_cluster_kernel<true>, p);
}

static int launch_approx_sym(const ApproxParams& p, cudaStream_t st) {
    static int occ_warp = 0, occ_sym = 0;
    if (p.B >= WPS_MIN_BATCH)
        return launch_persistent(approx_ndcg_warp_kernel, WPS_THREADS, WPS_WARPS, WPS_MIN_CTAS, p, st, occ_warp);
    return launch_persistent(approx_ndcg_sym_kernel, SYM_THREADS, 1, SYM_GRID_PER_SM, p, st, occ_sym);
}

}  // namespace ltr

namespace ltr {
int upload_disc_prefix_constant() {
    double acc = 0.0;
    float host[129];
    host[0] = 0.0f;
    for (int r = 1; r <= 128; ++r) {
        acc += 1.0 / log2(1.0 + (double)r);
        host[r] = (float)acc;
    }
    return (int)cudaMemcpyToSymbol(c_disc_prefix, host, sizeof(host));
}
}  // namespace ltr

LTR_API size_t ltr_approx_ndcg_scratch_bytes(int B, int n) {
    if (B <= 0 || n <= ltr::SYM_N) return 0;
    const size_t chunks = ((size_t)n + 127) / 128;
    return ((size_t)B * n + (size_t)B * chunks) * sizeof(float);
}

LTR_API int ltr_approx_ndcg_fwdbwd(const float* y_pred, const float* y_true, int B, int n, float alpha, float eps,
                                   float pad_value, const double* disc_prefix, int disc_len, float* slate_ndcg,
                                   float* loss, float* grad, void* scratch, size_t scratch_bytes, void* workspace,
                                   void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !slate_ndcg || !loss || !workspace || B <= 0 || n <= 0)
        return set_error(LTR_EINVAL, "ltr_approx_ndcg_fwdbwd: null pointer or non-positive B/n");
    if (n > ltr_max_slate_len()) return set_error(LTR_ETOOBIG, "ltr_approx_ndcg_fwdbwd: slate too long");
    ApproxParams p;
    p.s = y_pred;
    p.y = y_true;
    p.B = B;
    p.n = n;
    p.npad = ((n + 3) & ~3) + 4;
    p.alpha = alpha;
    p.eps = eps;
    p.pad = pad_value;
    p.inv_B = 1.0f / (float)B;
    p.disc_prefix = disc_prefix;
    p.disc_len = disc_prefix ? disc_len : 0;
    p.slate_out = slate_ndcg;
    p.loss_out = loss;
    p.grad = grad;
    p.ws = reinterpret_cast<unsigned*>(workspace);
    // the per-pair eps clamp moves pos_i (>= 1) by at most n*eps: below fp32 resolution for the default 1e-10
    p.clamp = ((double)n * (double)eps > 2e-7) ? 1 : 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    p.cbuf = reinterpret_cast<float*>(scratch);
    p.part = p.cbuf ? p.cbuf + (size_t)B * n : nullptr;
    p.rows_per_cta = 0;
    int rc;
    if (n <= SYM_N && !p.clamp && p.disc_prefix != nullptr && p.disc_len > SYM_N && B <= (1 << 22))
        rc = launch_approx_sym(p, st);                                 // symmetric pair schedule, 4 warps per slate
#ifndef APPROX_CLUSTER
#define APPROX_CLUSTER 1
#endif
    // One cluster of 8 CTAs per slate: lowest latency for a slate (the batch fits one resident wave of ~100 clusters),
    // and the best throughput once the pair work dwarfs the per-cluster overhead (n > 768).  Larger batches of
    // mid-length slates keep the one-CTA-per-slate kernel (measured 1024 x 512: 142 us against 247 us).
    else if (APPROX_CLUSTER && n > SYM_N && n <= CL_MAXN && !p.clamp && alpha > 0.f && p.disc_prefix != nullptr &&
             p.disc_len > n && (B <= 5 * device_sm_count() / CL_CTAS || n > 768))
        rc = launch_approx_cluster(p, st);
    else if (n > 256 && scratch != nullptr && scratch_bytes >= ltr_approx_ndcg_scratch_bytes(B, n) &&
             B < 8 * device_sm_count())
        rc = launch_approx_split(p, st);                               // few long slates: several CTAs per slate
    else rc = (n <= 512) ? launch_approx<128>(p, st) : launch_approx<256>(p, st);
    return rc ? set_cuda_error(rc, "ltr_approx_ndcg_fwdbwd") : 0;
}

// ---- host-buffer entry: H2D copies, the fused kernel, D2H copies, all on `stream` ---------------
LTR_API size_t ltr_approx_ndcg_host_scratch_bytes(int B, int n) {
    const size_t bn = (((size_t)B * n + 3) & ~(size_t)3);
    const size_t b4 = (((size_t)B + 3) & ~(size_t)3);
    return (3 * bn + b4 + 4) * sizeof(float);
}

LTR_API int ltr_approx_ndcg_fwdbwd_host(const float* h_y_pred, const float* h_y_true, int B, int n, float alpha,
                                        float eps, float pad_value, const double* disc_prefix, int disc_len,
                                        float* h_loss, float* h_grad, float* h_slate_ndcg, void* d_scratch,
                                        size_t scratch_bytes, void* workspace, void* stream) {
    using namespace ltr;
    if (!h_y_pred || !h_y_true || !h_loss || !d_scratch || B <= 0 || n <= 0)
        return set_error(LTR_EINVAL, "ltr_approx_ndcg_fwdbwd_host: null pointer or non-positive B/n");
    if (scratch_bytes < ltr_approx_ndcg_host_scratch_bytes(B, n))
        return set_error(LTR_EINVAL, "ltr_approx_ndcg_fwdbwd_host: scratch too small");
    const size_t bn = (((size_t)B * n + 3) & ~(size_t)3);
    const size_t b4 = (((size_t)B + 3) & ~(size_t)3);
    float* d_pred = reinterpret_cast<float*>(d_scratch);
    float* d_true = d_pred + bn;
    float* d_grad = d_true + bn;
    float* d_slate = d_grad + bn;
    float* d_loss = d_slate + b4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)B * n * sizeof(float);
    // LTR_HOST_ZERO_COPY=0 in the environment forces the staged copies (A/B runs of the host path: tools/pcie_probe_multi.py)
    static const bool zero_copy = [] {
        const char* e = getenv("LTR_HOST_ZERO_COPY");
        return !(e && e[0] == '0');
    }();
    if (zero_copy) {
        // Pinned, device-mapped host buffers (cudaHostAlloc / torch pin_memory under unified addressing): the kernel
        // reads the scores and labels and writes the gradient and the loss over PCIe itself -- one launch, no staging
        // copies, the transfers of a step overlap its pair arithmetic slate by slate.  Buffers that are not mapped
        // fall through to the staged path below.
        void *ms = nullptr, *my = nullptr, *ml = nullptr, *mg = nullptr, *mv = nullptr;
        bool mapped = cudaHostGetDevicePointer(&ms, const_cast<float*>(h_y_pred), 0) == cudaSuccess &&
                      cudaHostGetDevicePointer(&my, const_cast<float*>(h_y_true), 0) == cudaSuccess &&
                      cudaHostGetDevicePointer(&ml, h_loss, 0) == cudaSuccess;
        if (mapped && h_grad) mapped = cudaHostGetDevicePointer(&mg, h_grad, 0) == cudaSuccess;
        if (mapped && h_slate_ndcg) mapped = cudaHostGetDevicePointer(&mv, h_slate_ndcg, 0) == cudaSuccess;
        if (mapped)
            return ltr_approx_ndcg_fwdbwd(reinterpret_cast<const float*>(ms), reinterpret_cast<const float*>(my), B, n, alpha,
                                          eps, pad_value, disc_prefix, disc_len,
                                          mv ? reinterpret_cast<float*>(mv) : d_slate, reinterpret_cast<float*>(ml),
                                          reinterpret_cast<float*>(mg), nullptr, 0, workspace, stream);
        (void)cudaGetLastError();                          // not mapped: clear the sticky "invalid value"
    }
    LTR_CUDA(cudaMemcpyAsync(d_pred, h_y_pred, bytes, cudaMemcpyHostToDevice, st));
    LTR_CUDA(cudaMemcpyAsync(d_true, h_y_true, bytes, cudaMemcpyHostToDevice, st));
    int rc = ltr_approx_ndcg_fwdbwd(d_pred, d_true, B, n, alpha, eps, pad_value, disc_prefix, disc_len, d_slate, d_loss,
                                    h_grad ? d_grad : nullptr, nullptr, 0, workspace, stream);
    if (rc) return rc;
    LTR_CUDA(cudaMemcpyAsync(h_loss, d_loss, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_grad) LTR_CUDA(cudaMemcpyAsync(h_grad, d_grad, bytes, cudaMemcpyDeviceToHost, st));
    if (h_slate_ndcg) LTR_CUDA(cudaMemcpyAsync(h_slate_ndcg, d_slate, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    return 0;
}
