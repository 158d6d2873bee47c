// Pair weight schemes and the pair log-probability shared by the LambdaLoss kernels
// (losses/lambdaL.py:49-55, 96-127 of the reference).
#pragma once
#include "ltr_common.cuh"

namespace ltr {

enum LambdaScheme {
    SCHEME_NONE = 0,            // weighing_scheme=None and rankNet_scheme: w = 1
    SCHEME_NDCG1 = 1,           // G_i / D_i
    SCHEME_NDCG2 = 2,           // delta_|i-j| |G_i - G_j|
    SCHEME_LAMBDARANK = 3,      // |1/D_i - 1/D_j| |G_i - G_j|
    SCHEME_NDCG2PP = 4,         // mu * ndcgLoss2 + lambdaRank
    SCHEME_GTDIFF = 5,          // |y_i - y_j|
    SCHEME_GTDIFF_POW = 6,      // |y_i^2 - y_j^2|
};

template <int SCHEME>
__device__ __forceinline__ float pair_weight(int i, int j, const float* G, const float* invD, const float* dtab,
                                             const float* Y, float mu) {
    if (SCHEME == SCHEME_NONE) return 1.0f;
    if (SCHEME == SCHEME_NDCG1) return G[i] * invD[i];
    if (SCHEME == SCHEME_GTDIFF) return fabsf(fmaxf(Y[i], 0.f) - fmaxf(Y[j], 0.f));          // clamped labels (:37)
    if (SCHEME == SCHEME_GTDIFF_POW) {
        const float a = fmaxf(Y[i], 0.f), c = fmaxf(Y[j], 0.f);
        return fabsf(a * a - c * c);
    }
    const float dg = fabsf(G[i] - G[j]);
    const int gap = i > j ? i - j : j - i;
    if (SCHEME == SCHEME_NDCG2) return dtab[gap] * dg;
    const float lr = fabsf(invD[i] - invD[j]) * dg;
    if (SCHEME == SCHEME_LAMBDARANK) return lr;
    return mu * (dtab[gap] * dg) + lr;                           // SCHEME_NDCG2PP (lambdaL.py:115)
}

// log-probability of one pair and its derivative w.r.t. delta = s_hi - s_lo.
// PRECISE = false: MUFU ex2 / rcp / lg2 (the training loss, 3 SFU ops per pair).  PRECISE = true: exp2f / log2f at
// ~1 ulp, for the risk losses, whose effectiveness values are small differences of sums of these terms.
template <bool PRECISE = false>
__device__ __forceinline__ void pair_logprob(float delta, float w, float sig_scale, float log_eps2, float out_scale,
                                             float& l, float& dl) {
    // lambdaL.py:49-50: clamp to +-1e8, NaN (-inf - -inf) -> 0.  The clamp itself is subsumed by the exponent cap
    // below (2^126 and 2^-inf = 0 are reached long before |delta| = 1e8 for any sigma that is not denormal).
    delta = (delta != delta) ? 0.f : delta;
    const float x = fminf(-sig_scale * delta, 126.0f);
    const float e = PRECISE ? exp2f(x) : ex2_approx(x);           // exp(-sigma delta)
    const float q = 1.0f + e;
    const float sig = PRECISE ? 1.0f / q : rcp_approx(q);
    float l2 = PRECISE ? -log2f(q) : -lg2_approx(q);              // log2 sigmoid
    bool open = l2 >= log_eps2;                                   // inner clamp(min=eps) inactive
    l2 = fmaxf(l2, log_eps2);
    float lw = w * l2;                                            // log2 of (.)^w
    open = open && (lw >= log_eps2);                              // outer clamp inactive
    lw = fmaxf(lw, log_eps2);
    l = lw * out_scale;                                           // out_scale = 1 (binary) or ln 2 (natural)
    // d log_b(sig^w) / d delta = w sigma (1 - sig) / ln b;  1 - sig = e * sig
    dl = open ? w * (e * sig) : 0.f;
}

}  // namespace ltr
