// Host-side helpers shared by the C-ABI translation units (status codes, error text, device info).
#pragma once
#include <cuda_runtime.h>

#include "../../include/ltr_b200.h"

namespace ltr {

int set_error(int code, const char* msg);            // records msg for ltr_last_error(), returns code
int set_cuda_error(int cuda_code, const char* where);  // records "<where>: <cudaGetErrorString>"
int device_sm_count();                               // SM count of the current device (cached per device)
int device_max_optin_smem();                         // bytes of opt-in shared memory per CTA

}  // namespace ltr

#define LTR_CUDA(expr)                                                  \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) return ltr::set_cuda_error((int)_e, #expr); \
    } while (0)
