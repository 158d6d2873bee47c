// Pair tiles of the factorised approxNDCG form (losses/approxNDCG.py:27-53 of the reference): R docs of a lane against
// the R docs of the visiting column block, all in registers.  Shared by the K1 kernel (approx_ndcg_sym.cuh) and the fused
// ranker step (ranker_fused.cu).
#pragma once
#include "ltr_common.cuh"

namespace ltr {

// ---- pair tiles: R rows of this lane against the R docs of the visiting column block ---------------
// Columns are processed two at a time with packed fp32x2 arithmetic (FFMA2/FADD2/FMUL2 halve the issue
// slots of the FP32 work; the MUFU.RCP per pair stays scalar).  An odd last column is done in scalar.
//
// Rank pass: q = 1 + u_i v_j, r = rcp(q) = sigma_ij is added to BOTH the row and the column
// accumulator; the column side is turned into sum sigma_ji = (#rows met) - sum sigma_ij afterwards,
// so a pair costs 1 FFMA + 1 MUFU + 2 FADD.
template <int R>
__device__ __forceinline__ void rank_tile(const float (&urow)[R], const float (&vcol)[R], float h, float (&acc_r)[R],
                                          float (&acc_r_hi)[R], float (&acc_c)[R]) {
    const f32x2 one2 = pack2(1.0f, 1.0f);
    const f32x2 h2 = pack2(h, h);                // step weight (1/2 on the antipodal step) rides on the FMA
#pragma unroll
    for (int c = 0; c + 1 < R; c += 2) {
        const f32x2 v2 = pack2(vcol[c], vcol[c + 1]);
        f32x2 ac2 = pack2(acc_c[c], acc_c[c + 1]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const f32x2 q2 = fma2(pack2(urow[r], urow[r]), v2, one2);
            float q0, q1;
            unpack2(q2, q0, q1);
            const f32x2 r2 = pack2(rcp_approx(q0), rcp_approx(q1));
            ac2 = fma2(r2, h2, ac2);
            const f32x2 ar2 = fma2(r2, h2, pack2(acc_r[r], acc_r_hi[r]));
            unpack2(ar2, acc_r[r], acc_r_hi[r]);
        }
        unpack2(ac2, acc_c[c], acc_c[c + 1]);
    }
    if (R & 1) {
        constexpr int c = R - 1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float rr = rcp_approx(fmaf(urow[r], vcol[c], 1.0f));
            acc_r[r] = fmaf(rr, h, acc_r[r]);
            acc_c[c] = fmaf(rr, h, acc_c[c]);
        }
    }
}

// Gradient pass: w = sigma_ij * sigma_ji = r * (p r) (exact at both tails), e = w (c_j - c_i) goes to
// the row accumulator with + and to the column accumulator with -.
template <int R>
__device__ __forceinline__ void grad_tile(const float (&urow)[R], const float (&vcol)[R], const float (&crow)[R],
                                          const float (&ccol)[R], float (&acc_r)[R], float (&acc_r_hi)[R],
                                          float (&acc_c)[R]) {
    const f32x2 one2 = pack2(1.0f, 1.0f);
#pragma unroll
    for (int c = 0; c + 1 < R; c += 2) {
        const f32x2 v2 = pack2(vcol[c], vcol[c + 1]);
        const f32x2 cc2 = pack2(ccol[c], ccol[c + 1]);
        f32x2 ac2 = pack2(acc_c[c], acc_c[c + 1]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float cr = crow[r];
            const f32x2 p2 = mul2(pack2(urow[r], urow[r]), v2);
            const f32x2 q2 = add2(p2, one2);
            float q0, q1;
            unpack2(q2, q0, q1);
            const f32x2 r2 = pack2(rcp_approx(q0), rcp_approx(q1));
            const f32x2 w2 = mul2(r2, mul2(p2, r2));
            const f32x2 e2 = mul2(w2, sub2(cc2, pack2(cr, cr)));
            ac2 = sub2(ac2, e2);
            const f32x2 ar2 = add2(pack2(acc_r[r], acc_r_hi[r]), e2);
            unpack2(ar2, acc_r[r], acc_r_hi[r]);
        }
        unpack2(ac2, acc_c[c], acc_c[c + 1]);
    }
    if (R & 1) {
        constexpr int c = R - 1;
        const float cc = ccol[c];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float cr = crow[r];
            const float pp = urow[r] * vcol[c];
            const float rr = rcp_approx(pp + 1.0f);
            const float e = (rr * (pp * rr)) * (cc - cr);
            acc_r[r] += e;
            acc_c[c] -= e;
        }
    }
}

// R consecutive floats of a shared array at a multiple of R: one 64- / 128-bit access for R = 2 / 4 (a lane stride of 2 or 4
// words is a 2- / 4-way bank conflict with scalar accesses; R = 3 is conflict-free as it is)
template <int R>
__device__ __forceinline__ void lds_vec(const float* p, float (&out)[R]) {
    if (R == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        out[0] = t.x, out[1] = t.y, out[2] = t.z, out[3 % R] = t.w;
    } else if (R == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        out[0] = t.x, out[1] = t.y;
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = p[r];
    }
}
template <int R>
__device__ __forceinline__ void sts_vec(float* p, const float (&in)[R]) {
    if (R == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3 % R]);
    } else if (R == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(in[0], in[1]);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) p[r] = in[r];
    }
}

}  // namespace ltr
