// Measurement probes used by bench.py to obtain the roofline denominators on the box itself:
// a pure MUFU (ex2.approx) loop for the SFU peak and an FFMA loop for the FP32 peak.
// "the 16/clk/SM MUFU rate is architectural knowledge, not measured -- the
// builder must microbenchmark a pure ex2.approx loop on the box".
#include "ltr_common.cuh"
#include "ltr_host.h"

namespace ltr {

// each thread runs `iters` x 8 independent ex2.approx chains
__global__ void __launch_bounds__(256) probe_sfu_kernel(float* out, int iters) {
    float x0 = 0.001f * threadIdx.x, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    float x4 = x0 + 0.4f, x5 = x0 + 0.5f, x6 = x0 + 0.6f, x7 = x0 + 0.7f;
    for (int i = 0; i < iters; ++i) {
        x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
        x4 = ex2_approx(x4); x5 = ex2_approx(x5); x6 = ex2_approx(x6); x7 = ex2_approx(x7);
        // keep values bounded without extra SFU work: map back towards 0 with one FFMA each
        x0 = fmaf(x0, 1e-3f, -1e-3f); x1 = fmaf(x1, 1e-3f, -1e-3f); x2 = fmaf(x2, 1e-3f, -1e-3f);
        x3 = fmaf(x3, 1e-3f, -1e-3f); x4 = fmaf(x4, 1e-3f, -1e-3f); x5 = fmaf(x5, 1e-3f, -1e-3f);
        x6 = fmaf(x6, 1e-3f, -1e-3f); x7 = fmaf(x7, 1e-3f, -1e-3f);
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;   // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) probe_fma_kernel(float* out, int iters) {
    float x0 = 0.001f * threadIdx.x, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    float x4 = x0 + 0.4f, x5 = x0 + 0.5f, x6 = x0 + 0.6f, x7 = x0 + 0.7f;
    const float a = 0.999f, b = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;
}

}  // namespace ltr

// Launches grid = 8 CTAs/SM x 256 threads; SFU ops executed = grid*256*iters*8.
LTR_API int ltr_probe_sfu(float* out, int iters, long long* ops_out, void* stream) {
    if (!out || iters <= 0) return ltr::set_error(LTR_EINVAL, "ltr_probe_sfu: bad arguments");
    const int grid = ltr::device_sm_count() * 8;
    ltr::probe_sfu_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, iters);
    if (ops_out) *ops_out = (long long)grid * 256LL * iters * 8LL;
    int rc = (int)cudaGetLastError();
    return rc ? ltr::set_cuda_error(rc, "ltr_probe_sfu") : 0;
}

// FP32 lane-ops executed (FMA counted once) = grid*256*iters*32.
LTR_API int ltr_probe_fma(float* out, int iters, long long* ops_out, void* stream) {
    if (!out || iters <= 0) return ltr::set_error(LTR_EINVAL, "ltr_probe_fma: bad arguments");
    const int grid = ltr::device_sm_count() * 8;
    ltr::probe_fma_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, iters);
    if (ops_out) *ops_out = (long long)grid * 256LL * iters * 32LL;
    int rc = (int)cudaGetLastError();
    return rc ? ltr::set_cuda_error(rc, "ltr_probe_fma") : 0;
}
