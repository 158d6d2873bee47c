// K6: risk-sensitive losses (replaces losses/riskLosses/riskFunctions.py:4-33 and the ListNet-style
// family of losses/riskLosses/riskLosses.py:8-60,128-180,247-291).
//
//   K6a  risk_rows_kernel      per query: softmax(y_true), softmax(system) for the model and every baseline
//                              system -> one row of the query x system effectiveness matrix  (streaming, HBM-bound:
//                              4 n (2 + S) bytes read per query)
//   K6b  risk_stats1 / risk_stats2 / risk_final   the global coupling of zRisk / geoRisk / tRisk: a matrix-wide max
//                              (riskLosses.py:50), column sums S_i, row sums T_q, N, the z sums, and the analytic
//                              gradient d loss / d mat[q, 0].  They are three tiny single-CTA kernels so that, when the
//                              queries are sharded over GPUs, the two exchange rounds sit between them:
//                              the 16-double state block of each phase is SUM/MAX all-reduced over NCCL by the host.
//   K6c  risk_rows_bwd_kernel  per query: d loss / d y_pred = gq[b] * d mat[b, 0] / d y_pred (softmax backward)
//
// Only column 0 (the model) depends on y_pred; baselines and the ideal column are constants.
#include "ltr_host.h"
#include "stream_common.cuh"

namespace ltr {

// ---------------------------------------------------------------------------------------------------
// K6a
// ---------------------------------------------------------------------------------------------------
struct RiskRowsParams {
    const float* y_pred;     // [B, n]
    const float* y_true;     // [B, n]
    const float* y_base;     // [B, n, S] or null
    int B, n, S, lt, add_ideal, tmode;
    float* mat;              // [B, C], C = 1 + S + add_ideal
    const float* gq;         // backward: [B] d loss / d mat[b, 0]
    float* grad;             // backward: [B, n]
};

// effectiveness of one system against the label distribution pt (both softmaxed, padding lanes 0)
template <int ITEMS>
__device__ __forceinline__ float effectiveness(const RowRegs<ITEMS>& pt, const RowRegs<ITEMS>& ps, int lt, int tmode) {
    if (lt == 1) {                                     // sum_d (pt ps - pt^2)^2          riskLosses.py:16-18
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const float d = pt.v[k] * ps.v[k] - pt.v[k] * pt.v[k];
            a = fmaf(d, d, a);
        }
        return warp_sum(a);
    }
    if (lt == 3) {                                     // (sum pt ps - sum pt^2)^2        riskLosses.py:21-23
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) a = fmaf(pt.v[k], ps.v[k], a), c = fmaf(pt.v[k], pt.v[k], c);
        const float u = warp_sum(a) - warp_sum(c);
        return u * u;
    }
    // lt == 2: cosine similarity, eps = 1e-8.  ListNet family: cos(pt, ps) (:19-20); t-risk: cos(pt^2, pt ps) (:262)
    float xy = 0.f, xx = 0.f, yy = 0.f;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const float x = tmode ? pt.v[k] * pt.v[k] : pt.v[k];
        const float y = tmode ? pt.v[k] * ps.v[k] : ps.v[k];
        xy = fmaf(x, y, xy), xx = fmaf(x, x, xx), yy = fmaf(y, y, yy);
    }
    xy = warp_sum(xy), xx = warp_sum(xx), yy = warp_sum(yy);
    return xy / (fmaxf(sqrtf(xx), 1e-8f) * fmaxf(sqrtf(yy), 1e-8f));
}

template <int ITEMS>
__global__ void __launch_bounds__(STREAM_THREADS) risk_rows_kernel(const RiskRowsParams p) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * STREAM_WARPS;
    const float ninf = __int_as_float(0xff800000);
    const int C = 1 + p.S + p.add_ideal;
    for (int b = gwarp; b < p.B; b += nwarps) {
        RowRegs<ITEMS> pt, ps;
        pt.load(p.y_true + (size_t)b * p.n, p.n, lane, ninf);
        softmax_row(pt);
        ps.load(p.y_pred + (size_t)b * p.n, p.n, lane, ninf);
        softmax_row(ps);
        float v = effectiveness(pt, ps, p.lt, p.tmode);
        if (lane == 0) p.mat[(size_t)b * C] = v;
        for (int j = 0; j < p.S; ++j) {
            ps.load_strided(p.y_base + (size_t)b * p.n * p.S + j, p.n, p.S, lane, ninf);
            softmax_row(ps);
            v = effectiveness(pt, ps, p.lt, p.tmode);
            if (lane == 0) p.mat[(size_t)b * C + 1 + j] = v;
        }
        if (p.add_ideal) {                             // the ideal ranking as one more system  (:35-41)
            v = effectiveness(pt, pt, p.lt, p.tmode);
            if (lane == 0) p.mat[(size_t)b * C + 1 + p.S] = v;
        }
    }
}

// K6c: grad[b, :] = gq[b] * d mat[b, 0] / d y_pred[b, :]
template <int ITEMS>
__global__ void __launch_bounds__(STREAM_THREADS) risk_rows_bwd_kernel(const RiskRowsParams p) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * STREAM_WARPS;
    const float ninf = __int_as_float(0xff800000);
    for (int b = gwarp; b < p.B; b += nwarps) {
        RowRegs<ITEMS> pt, ps, g;
        pt.load(p.y_true + (size_t)b * p.n, p.n, lane, ninf);
        softmax_row(pt);
        ps.load(p.y_pred + (size_t)b * p.n, p.n, lane, ninf);
        softmax_row(ps);
        // g_d = d value / d ps_d
        if (p.lt == 1) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) g.v[k] = 2.0f * (pt.v[k] * ps.v[k] - pt.v[k] * pt.v[k]) * pt.v[k];
        } else if (p.lt == 3) {
            float a = 0.f, c = 0.f;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) a = fmaf(pt.v[k], ps.v[k], a), c = fmaf(pt.v[k], pt.v[k], c);
            const float u = warp_sum(a) - warp_sum(c);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) g.v[k] = 2.0f * u * pt.v[k];
        } else {
            float xy = 0.f, xx = 0.f, yy = 0.f;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const float x = p.tmode ? pt.v[k] * pt.v[k] : pt.v[k];
                const float y = p.tmode ? pt.v[k] * ps.v[k] : ps.v[k];
                xy = fmaf(x, y, xy), xx = fmaf(x, x, xx), yy = fmaf(y, y, yy);
            }
            xy = warp_sum(xy), xx = warp_sum(xx), yy = warp_sum(yy);
            const float nx = fmaxf(sqrtf(xx), 1e-8f), ny = fmaxf(sqrtf(yy), 1e-8f);
            const float c = xy / (nx * ny);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const float x = p.tmode ? pt.v[k] * pt.v[k] : pt.v[k];
                const float y = p.tmode ? pt.v[k] * ps.v[k] : ps.v[k];
                const float dy = x / (nx * ny) - c * y / (ny * ny);      // d cos / d y_d
                g.v[k] = p.tmode ? dy * pt.v[k] : dy;
            }
        }
        // softmax backward: d/d s_k = ps_k (g_k - sum_d g_d ps_d)
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) dot = fmaf(g.v[k], ps.v[k], dot);
        dot = warp_sum(dot);
        const float up = p.gq[b];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) g.v[k] = up * ps.v[k] * (g.v[k] - dot);
        g.store(p.grad + (size_t)b * p.n, p.n, lane);
    }
}

// ---------------------------------------------------------------------------------------------------
// K6b.  One CTA of 1024 threads walks the [Q, C] matrix (a few KB .. MB); all arithmetic in fp64.
// ---------------------------------------------------------------------------------------------------
constexpr int RISKF_THREADS = 1024;
constexpr int RISK_STATE = 16;           // doubles per state block

struct RiskFinalParams {
    const float* mat;        // [Q, C] local rows (raw effectiveness, before the max - mat inversion)
    int Q, C;
    int kind;                // 0 zRisk, 1 geoRisk, 2 tRisk
    int invert;              // mat <- max(mat) - mat     (riskLosses.py:49-50)
    int strategy;            // return_strategy 1 / 2 / 3
    int quirk;               // zRiskListnetLoss: strategy 2 is factor * z(-1) - z(0)   (:176)
    float alpha, negative;
    double* state1;          // [RISK_STATE] phase-1 state (max in [0], sums from [1])
    double* state2;          // [RISK_STATE] phase-2 state
    float* loss;             // [1]
    float* gq;               // [Q] d loss / d mat[q, 0]
    float* gmat;             // [Q, C] d loss / d mat (full), or null
};

template <int THREADS>
__device__ __forceinline__ double rf_block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += red[w];
    return t;
}
template <int THREADS>
__device__ __forceinline__ double rf_block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(LTR_FULL_MASK, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
    for (int w = 1; w < THREADS / 32; ++w) t = fmax(t, red[w]);
    return t;
}

// phase 1 (local): state1 = {max(mat), sum_q mat[q,0], sum_q mat[q,C-1], sum of all entries, Q,
//                            t-risk: sum_q d_q, 0...} where d_q = mat[q,1] - mat[q,0] needs no max at all.
// `q0`, `dq`: first row and row stride of this thread (one CTA: threadIdx.x, RISKF_THREADS); `out`: where the block sums
// go (state1 itself, or this CTA's slot of the cluster's partials)
template <int THREADS>
__device__ __forceinline__ void risk_stats1_body(const RiskFinalParams& p, int q0, int dq, double* out) {
    __shared__ double red[32];
    double mx = -1.0e300, s0 = 0.0, sl = 0.0, sa = 0.0;
    for (int q = q0; q < p.Q; q += dq) {
        const float* row = p.mat + (size_t)q * p.C;
        for (int i = 0; i < p.C; ++i) {
            const double x = (double)row[i];
            mx = fmax(mx, x);
            sa += x;
        }
        s0 += (double)row[0];
        sl += (double)row[p.C - 1];
    }
    mx = rf_block_max<THREADS>(mx, red);
    s0 = rf_block_sum<THREADS>(s0, red);
    sl = rf_block_sum<THREADS>(sl, red);
    sa = rf_block_sum<THREADS>(sa, red);
    if (threadIdx.x == 0) {
        out[0] = mx;
        out[1] = s0;
        out[2] = sl;
        out[3] = sa;
        out[4] = (double)p.Q;
    }
}

// Everything phase 2/3 need about the z-risk of column `col` given the GLOBAL phase-1 state.
struct ZCol {
    double S, N, Qg, M;
    int col, invert, C;
    float alpha;
    __device__ __forceinline__ double X(const float* row, int i) const {
        return invert ? M - (double)row[i] : (double)row[i];
    }
    __device__ __forceinline__ double T(const float* row) const {
        double t = 0.0;
        for (int i = 0; i < C; ++i) t += X(row, i);
        return t;
    }
    // z_q, c_q = 1 + alpha [z < 0], A_q = c_q dz/de, e_q
    __device__ __forceinline__ void eval(const float* row, double& z, double& c, double& A, double& e, double& t) const {
        t = T(row);
        e = S * t / N;
        const double x = X(row, col);
        const double se = sqrt(e);
        z = (x - e) / se;
        c = (z < 0.0) ? 1.0 + (double)alpha : 1.0;
        A = -c * (x + e) / (2.0 * e * se);
    }
};

__device__ __forceinline__ ZCol make_zcol(const RiskFinalParams& p, const double* g1, bool last) {
    ZCol zc;
    zc.M = g1[0];
    zc.Qg = g1[4];
    zc.C = p.C;
    zc.invert = p.invert;
    zc.alpha = p.alpha;
    zc.col = last ? p.C - 1 : 0;
    const double sraw = last ? g1[2] : g1[1];
    zc.S = p.invert ? zc.Qg * zc.M - sraw : sraw;
    zc.N = p.invert ? zc.Qg * (double)p.C * zc.M - g1[3] : g1[3];
    return zc;
}

// phase 2 (local, needs global state1): per used column {sum z c, sum A T, sum A e, sum c / sqrt(e), sum A}, count of
// entries equal to the global max;  t-risk: {sum d', sum d'^2} with d' = d (1 + alpha [d < 0]).
template <int THREADS>
__device__ __forceinline__ void risk_stats2_body(const RiskFinalParams& p, int q0, int dq, double* out) {
    __shared__ double red[32];
    const double* g1 = p.state1;
    double a[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double nmax = 0.0, td = 0.0, td2 = 0.0;
    if (p.kind == 2) {
        for (int q = q0; q < p.Q; q += dq) {
            const float* row = p.mat + (size_t)q * p.C;
            // ours - base after the (optional) inversion; the max cancels
            double d = p.invert ? (double)row[1] - (double)row[0] : (double)row[0] - (double)row[1];
            d = (d < 0.0) ? d * (1.0 + (double)p.alpha) : d;
            td += d;
            td2 += d * d;
        }
    } else {
        const ZCol z0 = make_zcol(p, g1, false), zl = make_zcol(p, g1, true);
        for (int q = q0; q < p.Q; q += dq) {
            const float* row = p.mat + (size_t)q * p.C;
            double z, c, A, e, t;
            z0.eval(row, z, c, A, e, t);
            a[0] += z * c; a[1] += A * t; a[2] += A * e; a[3] += c / sqrt(e); a[4] += A;
            if (p.strategy != 1) {
                zl.eval(row, z, c, A, e, t);
                a[5] += z * c; a[6] += A * t; a[7] += A * e; a[8] += c / sqrt(e); a[9] += A;
            }
            if (p.invert)
                for (int i = 0; i < p.C; ++i) nmax += ((double)row[i] == g1[0]) ? 1.0 : 0.0;
        }
    }
    for (int j = 0; j < 10; ++j) a[j] = rf_block_sum<THREADS>(a[j], red);
    nmax = rf_block_sum<THREADS>(nmax, red);
    td = rf_block_sum<THREADS>(td, red);
    td2 = rf_block_sum<THREADS>(td2, red);
    if (threadIdx.x == 0) {
        for (int j = 0; j < 10; ++j) out[j] = a[j];
        out[10] = nmax;
        out[11] = td;
        out[12] = td2;
    }
}

__device__ __forceinline__ double std_normal_cdf(double x) { return 0.5 * erfc(-x * 0.7071067811865476); }

// phase 3 (local, needs global state1 and state2): loss value and gq[q] = d loss / d mat[q, 0].
__device__ __forceinline__ void risk_final_body(const RiskFinalParams& p, int q0, int dq, bool write_loss) {
    const double* g1 = p.state1;
    const double* g2 = p.state2;
    const double f = (double)p.negative;
    const double Qg = g1[4];
    if (p.kind == 2) {
        // t-risk: factor * mean(d') / std(d') (unbiased), riskLosses.py:280-291
        const double mu = g2[11] / Qg;
        const double var = (g2[12] - Qg * mu * mu) / (Qg - 1.0);
        const double sd = sqrt(var);
        if (write_loss && threadIdx.x == 0) p.loss[0] = (float)(f * mu / sd);
        for (int q = q0; q < p.Q; q += dq) {
            const float* row = p.mat + (size_t)q * p.C;
            double d = p.invert ? (double)row[1] - (double)row[0] : (double)row[0] - (double)row[1];
            const double c = (d < 0.0) ? 1.0 + (double)p.alpha : 1.0;
            d *= c;
            // d(mu/sd)/d d'_q = 1/(Q sd) - mu (d'_q - mu) / ((Q-1) sd^3)
            const double gd = f * (1.0 / (Qg * sd) - mu * (d - mu) / ((Qg - 1.0) * var * sd));
            const double g0 = gd * c * (p.invert ? -1.0 : 1.0);
            if (p.gq) p.gq[q] = (float)g0;
            if (p.gmat) {
                for (int i = 0; i < p.C; ++i) p.gmat[(size_t)q * p.C + i] = (i == 0) ? (float)g0 : (i == 1) ? (float)(-g0) : 0.f;
            }
        }
        return;
    }
    const ZCol z0 = make_zcol(p, g1, false), zl = make_zcol(p, g1, true);
    // value of each used column and d value / d (z sum), d value / d S (direct)
    double val[2], dz[2], dS[2];
    for (int u = 0; u < 2; ++u) {
        const ZCol& zc = u ? zl : z0;
        const double zsum = g2[5 * u];
        if (p.kind == 0) {
            val[u] = zsum; dz[u] = 1.0; dS[u] = 0.0;
        } else {
            const double arg = zsum / Qg;
            const double cdf = std_normal_cdf(arg);
            const double geo = sqrt(zc.S / Qg * cdf);
            const double pdf = 0.3989422804014327 * exp(-0.5 * arg * arg);
            val[u] = geo;
            dz[u] = (zc.S / Qg) * pdf / Qg / (2.0 * geo);
            dS[u] = cdf / Qg / (2.0 * geo);
        }
    }
    double loss, coef[2];                                       // d loss / d val[u]
    if (p.strategy == 1) {
        loss = f * val[0]; coef[0] = f; coef[1] = 0.0;
    } else if (p.strategy == 2) {
        if (p.quirk) { loss = f * val[1] - val[0]; coef[1] = f; coef[0] = -1.0; }
        else { loss = f * (val[1] - val[0]); coef[1] = f; coef[0] = -f; }
    } else {
        const double d = val[1] - val[0];
        loss = f * d * d; coef[1] = 2.0 * f * d; coef[0] = -coef[1];
    }
    if (write_loss && threadIdx.x == 0) p.loss[0] = (float)loss;
    if (!p.gq && !p.gmat) return;
    // d loss / d X[q', i'] = sum_u coef_u { dz_u ( [i' == col_u] (c/sqrt(e) + sum(A T)/N) + A_q' S_u / N - sum(A e)/N )
    //                                       + dS_u [i' == col_u] }
    // The max - mat inversion adds: d loss / d mat = -d loss / d X, plus (sum of all d loss / d X) / #maxima on every
    // entry equal to the max (torch.max distributes the gradient evenly).
    double all = 0.0;                                           // sum over every entry of d loss / d X
    const int nu = (p.strategy == 1) ? 1 : 2;
    for (int u = 0; u < nu; ++u) {
        const ZCol& zc = u ? zl : z0;
        const double sAT = g2[5 * u + 1], sAe = g2[5 * u + 2], sce = g2[5 * u + 3], sA = g2[5 * u + 4];
        all += coef[u] * (dz[u] * (sce + Qg * sAT / zc.N + (double)p.C * sA * zc.S / zc.N - Qg * (double)p.C * sAe / zc.N) +
                          dS[u] * Qg);
    }
    const double share = p.invert ? all / g2[10] : 0.0;
    for (int q = q0; q < p.Q; q += dq) {
        const float* row = p.mat + (size_t)q * p.C;
        double own[2], shared[2];
        for (int u = 0; u < nu; ++u) {
            const ZCol& zc = u ? zl : z0;
            double z, c, A, e, t;
            zc.eval(row, z, c, A, e, t);
            const double sAT = g2[5 * u + 1], sAe = g2[5 * u + 2];
            shared[u] = coef[u] * dz[u] * (A * zc.S / zc.N - sAe / zc.N);   // reaches every column through T_q and N
            own[u] = coef[u] * (dz[u] * (c / sqrt(e) + sAT / zc.N) + dS[u]); // only the column the risk is taken of
        }
        for (int i = 0; i < p.C; ++i) {
            if (i > 0 && !p.gmat) break;
            double g = 0.0;                                     // d loss / d X[q, i]
            for (int u = 0; u < nu; ++u) g += shared[u] + (((u ? zl.col : z0.col) == i) ? own[u] : 0.0);
            if (p.invert) g = -g + (((double)row[i] == g1[0]) ? share : 0.0);
            if (i == 0 && p.gq) p.gq[q] = (float)g;
            if (p.gmat) p.gmat[(size_t)q * p.C + i] = (float)g;
        }
    }
}

__global__ void __launch_bounds__(RISKF_THREADS) risk_stats1_kernel(const RiskFinalParams p) {
    pdl_prologue();
    risk_stats1_body<RISKF_THREADS>(p, threadIdx.x, RISKF_THREADS, p.state1);
}
__global__ void __launch_bounds__(RISKF_THREADS) risk_stats2_kernel(const RiskFinalParams p) {
    pdl_prologue();
    risk_stats2_body<RISKF_THREADS>(p, threadIdx.x, RISKF_THREADS, p.state2);
}
__global__ void __launch_bounds__(RISKF_THREADS) risk_final_kernel(const RiskFinalParams p) {
    pdl_prologue();
    risk_final_body(p, threadIdx.x, RISKF_THREADS, true);
}

// One GPU, no exchange between the phases: three launches of RISK_MC_BLOCKS small CTAs (one row per thread at 8192
// rows) instead of one CTA that does all the fp64 divisions and square roots of the matrix on a single SM.  A phase
// leaves its block sums in `part` (global scratch, [phase][block][RISK_STATE]); EVERY block of the next phase adds
// them up in block order into its own shared-memory copy of the state (deterministic, no counters, no atomics:
// the kernel boundary is the barrier), block 0 also publishes the state blocks.
constexpr int RISK_MC_BLOCKS = 64;
constexpr int RISK_MC_THREADS = 128;
struct RiskMcParams {
    RiskFinalParams f;
    double* part;            // [2][RISK_MC_BLOCKS][RISK_STATE]
};
__device__ __forceinline__ void risk_mc_gather(const double* part, int phase, double* dst, bool is_max0) {
    // thread j < RISK_STATE adds entry j of every block, in block order
    if (threadIdx.x < RISK_STATE) {
        const double* src = part + (size_t)phase * RISK_MC_BLOCKS * RISK_STATE + threadIdx.x;
        double t = src[0];
        for (int b = 1; b < RISK_MC_BLOCKS; ++b) {
            const double v = src[(size_t)b * RISK_STATE];
            t = (is_max0 && threadIdx.x == 0) ? fmax(t, v) : t + v;
        }
        dst[threadIdx.x] = t;
    }
}
template <int PHASE>
__global__ void __launch_bounds__(RISK_MC_THREADS) risk_mc_kernel(const RiskMcParams mp) {
    __shared__ double st1[RISK_STATE], st2[RISK_STATE];
    RiskFinalParams p = mp.f;
    const int q0 = (int)blockIdx.x * RISK_MC_THREADS + (int)threadIdx.x, dq = RISK_MC_BLOCKS * RISK_MC_THREADS;
    pdl_prologue();
    if (PHASE >= 2) {
        risk_mc_gather(mp.part, 0, st1, true);
        __syncthreads();
        if (threadIdx.x == 0) st1[4] = (double)p.Q;              // (every block reported Q: not a sum)
        if (PHASE == 3) risk_mc_gather(mp.part, 1, st2, false);
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x < RISK_STATE) {
            if (PHASE == 2) p.state1[threadIdx.x] = st1[threadIdx.x];
            else p.state2[threadIdx.x] = st2[threadIdx.x];
        }
        p.state1 = st1;
        p.state2 = st2;
    }
    double* mine = mp.part + ((size_t)(PHASE - 1) * RISK_MC_BLOCKS + blockIdx.x) * RISK_STATE;
    if (PHASE == 1) risk_stats1_body<RISK_MC_THREADS>(p, q0, dq, mine);
    else if (PHASE == 2) risk_stats2_body<RISK_MC_THREADS>(p, q0, dq, mine);
    else risk_final_body(p, q0, dq, blockIdx.x == 0);
}

// getGeoRiskDefault (utils/metrics.py:8-45): GeoRisk of EVERY system column of an fp64 query x system matrix, fp64
// throughout, with the reference's guard "expected value 0 -> z = 0" (:31-34) -- the driver builds this matrix with
// no_relevant=False, so all-zero query rows occur (computeMetrics.py:94-97).  One CTA per column.
__global__ void __launch_bounds__(RISKF_THREADS) georisk_default_kernel(const double* __restrict__ mat, int Q, int C,
                                                                        double alpha, double* __restrict__ out) {
    __shared__ double red[32];
    const int col = blockIdx.x;
    double n_all = 0.0, s_col = 0.0;
    for (int q = threadIdx.x; q < Q; q += RISKF_THREADS) {
        const double* row = mat + (size_t)q * C;
        double t = 0.0;
        for (int i = 0; i < C; ++i) t += row[i];
        n_all += t;
        s_col += row[col];
    }
    const double N = rf_block_sum<RISKF_THREADS>(n_all, red);
    const double S = rf_block_sum<RISKF_THREADS>(s_col, red);
    double zs = 0.0;
    for (int q = threadIdx.x; q < Q; q += RISKF_THREADS) {
        const double* row = mat + (size_t)q * C;
        double t = 0.0;
        for (int i = 0; i < C; ++i) t += row[i];
        const double e = S * (t / N);                            // :29
        const double d = row[col] - e;
        double z = (e != 0.0) ? d / sqrt(e) : 0.0;               // :31-34
        if (d < 0.0) z *= 1.0 + alpha;                           // :35-36
        zs += z;
    }
    const double Z = rf_block_sum<RISKF_THREADS>(zs, red);
    if (threadIdx.x == 0) out[col] = sqrt((S / (double)Q) * std_normal_cdf(Z / (double)Q));   // :41-43
}

static int risk_grid(int B) {
    long long g = ((long long)B + STREAM_WARPS - 1) / STREAM_WARPS;
    const long long cap = (long long)device_sm_count() * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

template <bool BWD>
static int launch_rows(const RiskRowsParams& p, cudaStream_t st) {
    const int grid = risk_grid(p.B);
#define LTR_ROWS(ITEMS) \
    return BWD ? launch_pdl(risk_rows_bwd_kernel<ITEMS>, grid, STREAM_THREADS, 0, st, p) \
               : launch_pdl(risk_rows_kernel<ITEMS>, grid, STREAM_THREADS, 0, st, p)
    if (p.n <= 128) { LTR_ROWS(4); }
    else if (p.n <= 256) { LTR_ROWS(8); }
    else if (p.n <= 512) { LTR_ROWS(16); }
    else { LTR_ROWS(32); }
#undef LTR_ROWS
}

}  // namespace ltr

LTR_API int ltr_risk_listnet_rows(const float* y_pred, const float* y_true, const float* y_baselines, int B, int n, int S,
                                  int transformation, int add_ideal, int t_variant, float* mat, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !mat || B <= 0 || n <= 0 || S < 0 || (S > 0 && !y_baselines))
        return set_error(LTR_EINVAL, "ltr_risk_listnet_rows: bad arguments");
    if (transformation < 1 || transformation > 3) return set_error(LTR_EINVAL, "ltr_risk_listnet_rows: transformation must be 1..3");
    if (n > 1024) return set_error(LTR_ETOOBIG, "ltr_risk_listnet_rows: slate longer than 1024");
    RiskRowsParams p{y_pred, y_true, y_baselines, B, n, S, transformation, add_ideal ? 1 : 0, t_variant ? 1 : 0, mat,
                     nullptr, nullptr};
    const int rc = launch_rows<false>(p, reinterpret_cast<cudaStream_t>(stream));
    return rc ? set_cuda_error(rc, "ltr_risk_listnet_rows") : 0;
}

LTR_API int ltr_risk_listnet_rows_bwd(const float* y_pred, const float* y_true, int B, int n, int transformation,
                                      int t_variant, const float* gq, float* grad, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !gq || !grad || B <= 0 || n <= 0)
        return set_error(LTR_EINVAL, "ltr_risk_listnet_rows_bwd: bad arguments");
    if (transformation < 1 || transformation > 3) return set_error(LTR_EINVAL, "ltr_risk_listnet_rows_bwd: transformation must be 1..3");
    if (n > 1024) return set_error(LTR_ETOOBIG, "ltr_risk_listnet_rows_bwd: slate longer than 1024");
    RiskRowsParams p{y_pred, y_true, nullptr, B, n, 0, transformation, 0, t_variant ? 1 : 0, nullptr, gq, grad};
    const int rc = launch_rows<true>(p, reinterpret_cast<cudaStream_t>(stream));
    return rc ? set_cuda_error(rc, "ltr_risk_listnet_rows_bwd") : 0;
}

// phase = 1, 2, 3 (see K6b above); state1 / state2: RISK_STATE doubles each, on the device.
LTR_API int ltr_risk_phase(int phase, const float* mat, int Q, int C, int kind, int invert, int return_strategy,
                           int precedence_quirk, float alpha, float negative, double* state1, double* state2,
                           float* loss, float* gq, float* gmat, void* stream) {
    using namespace ltr;
    if (!mat || Q <= 0 || C <= 0 || !state1 || !state2 || kind < 0 || kind > 2 || phase < 1 || phase > 3)
        return set_error(LTR_EINVAL, "ltr_risk_phase: bad arguments");
    if (kind == 2 && C < 2) return set_error(LTR_EINVAL, "ltr_risk_phase: t-risk needs the model and a baseline column");
    if (phase == 3 && !loss) return set_error(LTR_EINVAL, "ltr_risk_phase: phase 3 needs the loss pointer");
    RiskFinalParams p{mat, Q, C, kind, invert, return_strategy, precedence_quirk, alpha, negative, state1, state2, loss, gq,
                      gmat};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int rc = phase == 1 ? launch_pdl(risk_stats1_kernel, 1, RISKF_THREADS, 0, st, p)
                 : phase == 2 ? launch_pdl(risk_stats2_kernel, 1, RISKF_THREADS, 0, st, p)
                              : launch_pdl(risk_final_kernel, 1, RISKF_THREADS, 0, st, p);
    return rc ? set_cuda_error(rc, "ltr_risk_phase") : 0;
}

// Single-GPU fused call: rows -> the three risk phases, 64 CTAs each -> backward rows: five dependent launches issued
// from one host call (each starts its ramp-up under the previous one's tail).
// state: 32 + 2 * 64 * 16 = 2080 device doubles of scratch.  Query-sharded callers use the separate
// entry points above and all-reduce the two state blocks between the phases.
LTR_API int ltr_risk_listnet_fwdbwd(const float* y_pred, const float* y_true, const float* y_baselines, int B, int n, int S,
                                    int transformation, int add_ideal, int kind, int invert, int return_strategy,
                                    int precedence_quirk, float alpha, float negative, float* mat, double* state,
                                    float* loss, float* gq, float* grad, void* stream) {
    using namespace ltr;
    if (!state || !loss || (grad && !gq)) return set_error(LTR_EINVAL, "ltr_risk_listnet_fwdbwd: bad arguments");
    const int t_variant = (kind == 2) ? 1 : 0;
    const int C = 1 + S + (add_ideal ? 1 : 0);
    int rc = ltr_risk_listnet_rows(y_pred, y_true, y_baselines, B, n, S, transformation, add_ideal, t_variant, mat, stream);
    if (!rc) {
        if (kind < 0 || kind > 2) return set_error(LTR_EINVAL, "ltr_risk_listnet_fwdbwd: bad kind");
        if (kind == 2 && C < 2) return set_error(LTR_EINVAL, "ltr_risk_listnet_fwdbwd: t-risk needs the model and a baseline column");
        RiskFinalParams fp{mat, B, C, kind, invert, return_strategy, precedence_quirk, alpha, negative, state, state + 16,
                           loss, gq, nullptr};
        RiskMcParams mp{fp, state + 32};
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        rc = launch_pdl(risk_mc_kernel<1>, RISK_MC_BLOCKS, RISK_MC_THREADS, 0, st, mp);
        if (!rc) rc = launch_pdl(risk_mc_kernel<2>, RISK_MC_BLOCKS, RISK_MC_THREADS, 0, st, mp);
        if (!rc) rc = launch_pdl(risk_mc_kernel<3>, RISK_MC_BLOCKS, RISK_MC_THREADS, 0, st, mp);
        if (rc) return set_cuda_error(rc, "ltr_risk_listnet_fwdbwd");
    }
    if (!rc && grad) rc = ltr_risk_listnet_rows_bwd(y_pred, y_true, B, n, transformation, t_variant, gq, grad, stream);
    return rc;
}

// getGeoRiskDefault: mat [Q, C] device doubles -> out [C] device doubles (GeoRisk of every column).
LTR_API int ltr_georisk_default_f64(const double* mat, int Q, int C, double alpha, double* out, void* stream) {
    using namespace ltr;
    if (!mat || !out || Q <= 0 || C <= 0) return set_error(LTR_EINVAL, "ltr_georisk_default_f64: bad arguments");
    georisk_default_kernel<<<C, RISKF_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mat, Q, C, alpha, out);
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_georisk_default_f64") : 0;
}
