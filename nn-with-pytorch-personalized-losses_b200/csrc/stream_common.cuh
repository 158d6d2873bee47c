// Helpers for the HBM-bound "one warp streams one slate" kernels (ListNet, risk-ListNet rows, ordinal).
//
// Work assignment is STATIC (slate = global warp id + k * total warps), every warp keeps its partial
// sums in fp64, every CTA writes one fp64 partial per output, and the last CTA to finish adds the
// per-CTA partials in index order: batch sums are bitwise reproducible and need no second launch.
#pragma once
#include "ltr_common.cuh"

namespace ltr {

constexpr int STREAM_THREADS = 256;          // 8 warps per CTA
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
constexpr int STREAM_MAX_OUT = 4;            // batch sums a streaming kernel may produce

// A slate row held in registers: element k of lane l is row[l + 32 k] (coalesced 128-byte requests).
template <int ITEMS>
struct RowRegs {
    float v[ITEMS];
    __device__ __forceinline__ void load(const float* row, int n, int lane, float fill) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = lane + 32 * k;
            v[k] = (i < n) ? ld_stream(row + i) : fill;
        }
    }
    __device__ __forceinline__ void load_strided(const float* row, int n, int stride, int lane, float fill) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = lane + 32 * k;
            v[k] = (i < n) ? __ldg(row + (size_t)i * stride) : fill;
        }
    }
    __device__ __forceinline__ void store(float* row, int n, int lane) const {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = lane + 32 * k;
            if (i < n) row[i] = v[k];
        }
    }
    __device__ __forceinline__ float max_all() const {
        float m = v[0];
#pragma unroll
        for (int k = 1; k < ITEMS; ++k) m = fmaxf(m, v[k]);
        return warp_max(m);
    }
    __device__ __forceinline__ float sum_all() const {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) s += v[k];
        return warp_sum(s);
    }
};

// softmax over a register row in place: v <- exp(v - max) / sum; returns log(sum) + max = logsumexp.
// Padding elements must hold -inf (they become 0).
template <int ITEMS>
__device__ __forceinline__ float softmax_row(RowRegs<ITEMS>& r) {
    const float m = r.max_all();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) r.v[k] = __expf(r.v[k] - m);
    const float s = r.sum_all();
    const float inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) r.v[k] *= inv;
    return m + __logf(s);
}

// CTA-level finish: `vals[j]` (j < nout) are this THREAD's fp64 partials.  partials: [gridDim.x][STREAM_MAX_OUT]
// doubles in global memory; counter: one unsigned, left at 0.  The last CTA writes out[j] = (float)(scale[j] * sum).
__device__ __forceinline__ void stream_finish(const double (&vals)[STREAM_MAX_OUT], int nout, double* partials,
                                              unsigned* counter, float* out, const double (&scale)[STREAM_MAX_OUT],
                                              double* out64) {
    __shared__ double sred[STREAM_WARPS][STREAM_MAX_OUT];
    __shared__ int is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = 0; j < nout; ++j) {
        const double w = warp_sum(vals[j]);
        if (lane == 0) sred[warp][j] = w;
    }
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < nout; ++j) {
            double t = 0.0;
            for (int w = 0; w < STREAM_WARPS; ++w) t += sred[w][j];
            partials[(size_t)blockIdx.x * STREAM_MAX_OUT + j] = t;
        }
        __threadfence();
        const unsigned done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int j = 0; j < nout; ++j) {
            double t = 0.0;
            for (int i = tid; i < (int)gridDim.x; i += STREAM_THREADS)
                t += __ldcg(partials + (size_t)i * STREAM_MAX_OUT + j);
            // fixed shape of the reduction tree -> deterministic
            t = warp_sum(t);
            __syncthreads();
            if (lane == 0) sred[warp][j] = t;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < STREAM_WARPS; ++w) tot += sred[w][j];
                if (out) out[j] = (float)(scale[j] * tot);
                if (out64) out64[j] = scale[j] * tot;
            }
        }
        if (tid == 0) {
            *counter = 0u;
            __threadfence();
        }
    }
}

}  // namespace ltr
