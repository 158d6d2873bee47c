// Shared device helpers for the B200 (sm_100a) listwise-LTR kernels.
// No torch types, no allocation: kernels get raw device pointers from the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef LTR_API
#define LTR_API extern "C" __attribute__((visibility("default")))
#endif

#define LTR_FULL_MASK 0xffffffffu

namespace ltr {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- MUFU wrappers (approximate, flush-to-zero; one SFU op each) -----------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 2^min(x, 126); a NaN exponent stays NaN (fminf alone would swallow it and hide a NaN score)
__device__ __forceinline__ float ex2_capped(float x) {
    return ex2_approx((x != x) ? x : fminf(x, 126.0f));
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- packed fp32x2 math (Blackwell FFMA2 / FADD2 / FMUL2) -------------------------------
struct f32x2 {
    unsigned long long bits;
};
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.bits) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v.bits));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits), "l"(c.bits));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
    return r;
}

// ---- warp / block reductions -------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LTR_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LTR_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(LTR_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(LTR_FULL_MASK, v, o));
    return v;
}

// Block-wide sum; `scratch` holds >= 32 floats. Result valid in every thread. Two barriers.
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    constexpr int W = THREADS / 32;
    v = warp_sum(v);
    if (W == 1) return v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) t += scratch[w];   // fixed order -> deterministic
    return t;
}

// ---- single-instruction warp reductions (REDUX) on order-preserving integer keys ---------------
__device__ __forceinline__ int float_order_key(float f) {
    const int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);          // monotone: a < b  <=>  key(a) < key(b)
}
__device__ __forceinline__ float float_from_key(int k) {
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_max_redux(float v) {
    return float_from_key(__reduce_max_sync(LTR_FULL_MASK, float_order_key(v)));
}
__device__ __forceinline__ float warp_min_redux(float v) {
    return float_from_key(__reduce_min_sync(LTR_FULL_MASK, float_order_key(v)));
}

// ---- mbarrier (shared::cta) producer/consumer primitives ----------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// named barrier among `count` threads (count a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- programmatic dependent launch (kernels started with ltr::launch_pdl) --------------------------------------
// Let the next launch on the stream become resident now; then wait until everything before this kernel on the
// stream is complete and visible (inputs may come from the previous kernel, the workspace from the previous call).
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- streaming global access -------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// 2^l - 1 exactly for small non-negative integer labels, exp2f otherwise (torch.pow(2, y) - 1).
__device__ __forceinline__ float gain_exp2(float l) {
    if (l >= 0.f && l <= 30.f && l == floorf(l)) return (float)(1u << (int)l) - 1.0f;
    return exp2f(l) - 1.0f;
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// In-shared-memory bitonic sort, DESCENDING by key, all threads of the block participate.
// `len` must be a power of two. Keys only (used for ideal-DCG label order, where ties are
// value-identical and so order-free).
// ---- "group" = the CTA (default) or one warp: the shared-memory sorts below run at either scope ----------------
template <bool WARP> __device__ __forceinline__ int grp_tid() { return WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
template <bool WARP> __device__ __forceinline__ int grp_size() { return WARP ? 32 : (int)blockDim.x; }
template <bool WARP> __device__ __forceinline__ void grp_sync() {
    if (WARP) __syncwarp();
    else __syncthreads();
}
template <bool WARP> __device__ __forceinline__ int grp_or(int v) {
    if (WARP) return __any_sync(0xffffffffu, v);
    return __syncthreads_or(v);
}

template <bool WARP = false>
__device__ __forceinline__ void block_bitonic_desc(float* keys, int len) {
#pragma unroll 1
    for (int k = 2; k <= len; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            grp_sync<WARP>();
            for (int t = grp_tid<WARP>(); t < (len >> 1); t += grp_size<WARP>()) {
                // index of the lower element of the t-th compare-exchange pair
                int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int hi = lo | j;
                bool desc = ((lo & k) == 0);
                float a = keys[lo], b = keys[hi];
                bool swap = desc ? (a < b) : (a > b);
                if (swap) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    grp_sync<WARP>();
}

}  // namespace ltr
