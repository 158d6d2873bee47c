// Small-tensor all-reduce (average) over the GPUs of one box through peer memory, without a collective library call:
// the gradient exchange that follows the fused ranker step under data parallelism (the reference's DDP loop,
// main_batch_execution.py:120-171, averages ~13 k gradient floats per step -- at that size a NCCL all-reduce is pure
// latency: 17 us on 2 GPUs, 35 us on 8, against a 90 us step).
//
// Every rank owns an exchange block in peer-mappable memory (cudaMalloc + CUDA IPC, like the risk losses' blocks):
//   float inbox[2][GX_MAX_WORLD][GX_SLOT_FLOATS]   one slot per (parity of the call, sending rank)
//   unsigned flag[2][GX_MAX_WORLD][GX_MAX_CTAS], seq, status, counter
// ONE launch per call, CTA b owning values [1024 b, 1024 b + 1024) on every rank; call number q is kept in the block
// (nothing per call comes from the host, so the launch can sit in a CUDA graph):
//   push    CTA b stores its values into slot [q & 1][rank] of EVERY rank's inbox (plain stores over NVLink), fences, and
//           release-stores q into flag[q & 1][rank][b] of every rank;
//   gather  CTA b waits (bounded) until its own flags [q & 1][*][b] carry q, adds the world's slots in rank order -- every
//           rank the same sum, bit for bit -- scales by 1 / world and writes the result back over its values.
// No CTA waits for another CTA of its own grid, only for the same CTA of the peers' grids, which run on other GPUs.
// Two parities: a slot is rewritten by call q + 2, which a rank can only start after its own call q + 1 has completed,
// which waited for every peer's push of q + 1, which every peer launched after its own call q had completed (stream order).
// A peer that never arrives (its process died) leaves a status word and NaNs instead of a hang.
#include <string.h>

#include "ltr_common.cuh"
#include "ltr_host.h"

namespace ltr {

constexpr int GX_MAX_WORLD = 8;
constexpr int GX_MAX_SEG = 8;
constexpr int GX_SLOT_FLOATS = 16384;                                   // 64 KB per (parity, rank)
constexpr int GX_CHUNK = 1024, GX_MAX_CTAS = GX_SLOT_FLOATS / GX_CHUNK;           // values per CTA (4 per thread)
constexpr size_t GX_WORDS_AT = (size_t)2 * GX_MAX_WORLD * GX_SLOT_FLOATS;      // (in floats)
constexpr int GX_W_SEQ = 2 * GX_MAX_WORLD * GX_MAX_CTAS, GX_W_STATUS = GX_W_SEQ + 1, GX_W_COUNTER = GX_W_SEQ + 2;
constexpr size_t GX_BYTES = GX_WORDS_AT * 4 + (GX_W_SEQ + 64) * 4;
constexpr unsigned long long GX_SPIN_NS = 4000000000ull;                // give up on a peer after 4 s

struct GxSegs {
    float* ptr[GX_MAX_SEG];
    int off[GX_MAX_SEG + 1];                                            // prefix sums of the segment lengths
    int nseg;
};
struct GxComm {
    float* xchg;
    float* peer[GX_MAX_WORLD];
    int rank, world;
};

__device__ __forceinline__ unsigned* gx_words(float* x) { return reinterpret_cast<unsigned*>(x + GX_WORDS_AT); }
__device__ __forceinline__ float* gx_slot(float* x, int par, int r) { return x + ((size_t)par * GX_MAX_WORLD + r) * GX_SLOT_FLOATS; }
__device__ __forceinline__ float* gx_value(const GxSegs& s, int i) {
    int k = 0;
#pragma unroll
    for (int j = 1; j < GX_MAX_SEG; ++j) k += (j < s.nseg && i >= s.off[j]) ? 1 : 0;
    return s.ptr[k] + (i - s.off[k]);
}

__global__ void __launch_bounds__(256) gx_allreduce_kernel(const GxSegs s, const GxComm c) {
    pdl_prologue();
    unsigned* words = gx_words(c.xchg);
    unsigned seq = __ldcg(words + GX_W_SEQ) + 1u;                       // every CTA reads it before the last one advances it
    if (seq == 0u) seq = 1u;                                            // (the flags start at 0)
    const int par = (int)(seq & 1u), total = s.off[s.nseg], b = blockIdx.x, tid = threadIdx.x;
    // push
#pragma unroll
    for (int k = 0; k < GX_CHUNK / 256; ++k) {
        const int i = b * GX_CHUNK + k * 256 + tid;
        if (i < total) {
            const float v = *gx_value(s, i);
            for (int r = 0; r < c.world; ++r)
                asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(gx_slot(c.peer[r], par, c.rank) + i), "f"(v) : "memory");
        }
    }
    __syncthreads();                                                    // (the release below is cumulative over the CTA's stores)
    if (tid < c.world)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(gx_words(c.peer[tid]) + (par * GX_MAX_WORLD + c.rank) * GX_MAX_CTAS + b),
                     "r"(seq) : "memory");
    // gather
    if (tid < c.world) {
        const unsigned* f = words + (par * GX_MAX_WORLD + tid) * GX_MAX_CTAS + b;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v == seq) break;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > GX_SPIN_NS) {                                 // a peer never arrived: poison the result, say so
                atomicExch(words + GX_W_STATUS, 1u);
                break;
            }
        }
    }
    __syncthreads();
    const bool dead = __ldcg(words + GX_W_STATUS) != 0u;
    const float inv = 1.0f / (float)c.world;
#pragma unroll
    for (int k = 0; k < GX_CHUNK / 256; ++k) {
        const int i = b * GX_CHUNK + k * 256 + tid;
        if (i < total) {
            float t = __ldcg(gx_slot(c.xchg, par, 0) + i);
            for (int r = 1; r < c.world; ++r) t += __ldcg(gx_slot(c.xchg, par, r) + i);      // rank order: the same bits on every rank
            *gx_value(s, i) = dead ? __int_as_float(0x7fc00000) : t * inv;
        }
    }
    // the last CTA out advances the call number (every CTA has read it by then)
    if (tid == 0 && atomicAdd(words + GX_W_COUNTER, 1u) == gridDim.x - 1) {
        words[GX_W_COUNTER] = 0u;
        words[GX_W_SEQ] = seq;
    }
}

}  // namespace ltr

LTR_API size_t ltr_peer_xchg_bytes(void) { return ltr::GX_BYTES; }
LTR_API int ltr_peer_xchg_max_floats(void) { return ltr::GX_SLOT_FLOATS; }

LTR_API int ltr_peer_xchg_alloc(void** xchg) {
    using namespace ltr;
    if (!xchg) return set_error(LTR_EINVAL, "ltr_peer_xchg_alloc: null pointer");
    LTR_CUDA(cudaMalloc(xchg, GX_BYTES));
    LTR_CUDA(cudaMemset(*xchg, 0, GX_BYTES));
    LTR_CUDA(cudaDeviceSynchronize());
    return 0;
}
LTR_API int ltr_peer_xchg_free(void* xchg) {
    using namespace ltr;
    LTR_CUDA(cudaFree(xchg));
    return 0;
}
LTR_API int ltr_peer_xchg_status(const void* xchg, int* status) {
    using namespace ltr;
    if (!xchg || !status) return set_error(LTR_EINVAL, "ltr_peer_xchg_status: null pointer");
    unsigned v = 0;
    LTR_CUDA(cudaMemcpy(&v, reinterpret_cast<const unsigned*>(static_cast<const float*>(xchg) + GX_WORDS_AT) + GX_W_STATUS, sizeof(v),
                        cudaMemcpyDeviceToHost));
    *status = (int)v;
    return 0;
}

LTR_API int ltr_peer_allreduce_avg(float* const* ptrs, const int* counts, int nseg, void* xchg, void* const* peer_xchg, int rank,
                                   int world, void* stream) {
    using namespace ltr;
    if (!ptrs || !counts || nseg < 1 || nseg > GX_MAX_SEG || !xchg || !peer_xchg || world < 1 || world > GX_MAX_WORLD || rank < 0 ||
        rank >= world)
        return set_error(LTR_EINVAL, "ltr_peer_allreduce_avg: bad arguments (1..8 segments, 1..8 ranks)");
    GxSegs s;
    memset(&s, 0, sizeof(s));
    s.nseg = nseg;
    int total = 0;
    for (int k = 0; k < nseg; ++k) {
        if (!ptrs[k] || counts[k] <= 0) return set_error(LTR_EINVAL, "ltr_peer_allreduce_avg: empty segment");
        s.ptr[k] = ptrs[k];
        s.off[k] = total;
        total += counts[k];
    }
    s.off[nseg] = total;
    for (int k = nseg + 1; k <= GX_MAX_SEG; ++k) s.off[k] = total;
    if (total > GX_SLOT_FLOATS) return set_error(LTR_ETOOBIG, "ltr_peer_allreduce_avg: more than 16,384 floats (use NCCL for large tensors)");
    GxComm c;
    memset(&c, 0, sizeof(c));
    c.xchg = static_cast<float*>(xchg);
    c.rank = rank;
    c.world = world;
    for (int r = 0; r < world; ++r) {
        if (!peer_xchg[r]) return set_error(LTR_EINVAL, "ltr_peer_allreduce_avg: null peer block");
        c.peer[r] = static_cast<float*>(peer_xchg[r]);
    }
    if (c.peer[rank] != c.xchg) return set_error(LTR_EINVAL, "ltr_peer_allreduce_avg: peer_xchg[rank] must be this rank's own block");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((total + GX_CHUNK - 1) / GX_CHUNK);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int rc = (int)cudaLaunchKernelEx(&cfg, gx_allreduce_kernel, s, c);
    if (rc) return set_cuda_error(rc, "ltr_peer_allreduce_avg");
    return 0;
}
