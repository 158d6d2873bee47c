// Library identification, error reporting, device queries and the discount prefix table.
#include <stdio.h>
#include <string.h>

#include "ltr_common.cuh"
#include "ltr_host.h"

namespace ltr {

static thread_local char g_err[256] = "";

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int set_cuda_error(int cuda_code, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString((cudaError_t)cuda_code));
    return cuda_code;
}

static int g_sm[64];
static int g_smem[64];
static void query_device(int dev) {
    cudaDeviceGetAttribute(&g_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}
int device_sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_sm[dev] == 0) query_device(dev);
    return g_sm[dev] > 0 ? g_sm[dev] : 148;
}
int device_max_optin_smem() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_smem[dev] == 0) query_device(dev);
    return g_smem[dev] > 0 ? g_smem[dev] : 227 * 1024;
}

// P[r] = sum_{q=1..r} 1/log2(1+q).  One block: terms in parallel, then a sequential fp64 prefix
// (len <= ~16k, built once per device).
__global__ void disc_prefix_kernel(double* P, int len) {
    for (int r = threadIdx.x; r < len; r += blockDim.x) P[r] = (r == 0) ? 0.0 : 1.0 / log2(1.0 + (double)r);
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int r = 0; r < len; ++r) {
            acc += P[r];
            P[r] = acc;
        }
    }
}

}  // namespace ltr

LTR_API const char* ltr_version(void) { return "ltr_b200 0.1.0 sm_100a"; }

LTR_API const char* ltr_last_error(void) { return ltr::g_err; }

LTR_API int ltr_max_slate_len(void) {
    // pairwise kernels keep up to 6 fp32 arrays of the slate in shared memory
    int n = ltr::device_max_optin_smem() / (6 * 4) - 16;
    return n > 8192 ? 8192 : n;
}

LTR_API int ltr_fill_disc_prefix(double* disc_prefix, int len, void* stream) {
    if (!disc_prefix || len <= 0) return ltr::set_error(LTR_EINVAL, "ltr_fill_disc_prefix: bad arguments");
    ltr::disc_prefix_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(disc_prefix, len);
    int rc = (int)cudaGetLastError();
    if (!rc) rc = ltr::upload_disc_prefix_constant();       // the approxNDCG fast path reads P[0..128] from the constant bank
    return rc ? ltr::set_cuda_error(rc, "ltr_fill_disc_prefix") : 0;
}
