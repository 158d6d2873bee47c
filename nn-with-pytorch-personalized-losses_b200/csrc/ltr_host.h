// Host-side helpers shared by the C-ABI translation units (status codes, error text, device info).
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/ltr_b200.h"

namespace ltr {

int set_error(int code, const char* msg);            // records msg for ltr_last_error(), returns code
int set_cuda_error(int cuda_code, const char* where);  // records "<where>: <cudaGetErrorString>"
int device_sm_count();                               // SM count of the current device (cached per device)
int device_max_optin_smem();                         // bytes of opt-in shared memory per CTA
int upload_disc_prefix_constant();                   // P[0..128] -> the constant bank of the current device

// ---- layout of the caller-provided workspace (LTR_WORKSPACE_BYTES, zero-filled once) -------------------------
//   u32 [0..4]    approxNDCG: next-slate counter, finished-CTA counter, 64-bit fixed-point loss sum, non-finite flag
//   u32 [16]      streaming kernels: finished-CTA counter
//   u32 [17..19]  LambdaLoss: exited-CTA counter, next-slate counter; schedule kernel: finished-CTA counter
//   byte 1024 ..  LambdaLoss schedule (batches of 296..4096 slates): keys, order, claim flags, arrival counters (rewritten per call)
//   f32 [32..35]  ordinal: BCE sum, number of valid docs, #targets < 0, #inputs outside [0, 1]
//   u32 [20..22]  risk losses: finished-CTA counters of phases A and B (three-launch form); cooperative kernel: generation
//   byte 1024 ..  streaming kernels: per-CTA fp64 partial sums, [grid <= 2048][4]
//   byte 256 ..   risk losses, cooperative kernel: the combined states of the two rounds, [2][16] stamped slots of 16 bytes
//   byte 768 ..   risk losses, cooperative kernel: 7 globaltimer stamps of CTA 0 (diagnostics)
//   byte 66560 .. risk losses: per-CTA fp64 partials of phase A [grid <= 1184][4], then of phase B [grid <= 128][16]
//                 (cooperative kernel: the same two regions hold stamped 16-byte slots, [grid <= 592][4] and [grid <= 592][16])
constexpr int LTR_WS_STREAM_COUNTER = 16;
constexpr int LTR_WS_RISK_COUNTERS = 20;
constexpr int LTR_WS_RISK_MAX_GRID = 1184;
constexpr size_t LTR_WS_RISK_PART1 = 1024 + 2048 * 4 * 8;
constexpr size_t LTR_WS_RISK_PART2 = LTR_WS_RISK_PART1 + (size_t)LTR_WS_RISK_MAX_GRID * 4 * 8;
constexpr int LTR_WS_RISK_COOP_MAX_GRID = 592;           // cooperative kernel: 4 CTAs per SM at most
static_assert(LTR_WS_RISK_PART2 + (size_t)LTR_WS_RISK_COOP_MAX_GRID * 16 * 16 <= LTR_WORKSPACE_BYTES, "workspace layout overflows");
static_assert((size_t)LTR_WS_RISK_COOP_MAX_GRID * 4 * 16 <= LTR_WS_RISK_PART2 - LTR_WS_RISK_PART1, "workspace layout overflows");
inline double* ws_risk_part1(void* ws) { return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + LTR_WS_RISK_PART1); }
inline double* ws_risk_part2(void* ws) { return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + LTR_WS_RISK_PART2); }
constexpr int LTR_WS_LAMBDA_COUNTER = 18;             // LambdaLoss: next-slate counter
constexpr int LTR_WS_LAMBDA_FINISHED = 17;            // LambdaLoss: exited-CTA counter (the last one reduces the slates' sums)
constexpr int LTR_WS_LAMBDA_ORDER_DONE = 19;          // LambdaLoss schedule: finished-CTA counter
constexpr int LTR_WS_ORDINAL_SUMS = 32;
inline unsigned* ws_counter(void* ws, int idx) { return reinterpret_cast<unsigned*>(ws) + idx; }
inline float* ws_floats(void* ws, int idx) { return reinterpret_cast<float*>(ws) + idx; }
inline double* ws_partials(void* ws) { return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 1024); }

// Launch with programmatic stream serialization: the grid may become resident while the previous kernel on the
// stream drains.  The kernel must execute `pdl_prologue()` before its first global-memory access.
template <typename Params>
inline int launch_pdl(void (*kernel)(Params), int grid, int threads, size_t smem, cudaStream_t st, const Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, p);
}

}  // namespace ltr

#define LTR_CUDA(expr)                                                  \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) return ltr::set_cuda_error((int)_e, #expr); \
    } while (0)
