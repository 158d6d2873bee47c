// Fused ranker tail + listwise loss (BASELINE configs[1]): TripleLayerNet forward, approxNDCG forward +
// backward and the ranker's weight gradients in ONE pass over the features.
//
//   reference step (main_batch_execution.py:128-170 with architeture/tripleLayer.py:5-17 and losses/approxNDCG.py:7-53)
//       h1 = l1(x)            136 -> 64, NO activation (the sigmoid is commented out, tripleLayer.py:13-14)
//       a  = sigmoid(l2(h1))   64 -> 32
//       s  = l3(a)             32 -> 1
//       loss = approxNDCGLoss(s, y);  loss.backward()
//   l1 and l2 compose linearly, so with Wc = W2 W1 [32, F] and bc = W2 b1 + b2 the forward is
//       z = x Wc^T + bc,  a = sigmoid(z),  s = a . w3 + b3
//   and every weight gradient follows from four small quantities (the host-side fold, ltr_b200/architeture):
//       dWc = dZ^T X,  dbc = sum dZ,  dw3 = A^T ds,  db3 = sum ds,      dZ = (ds w3^T) * a (1 - a)
//       dW1 = W2^T dWc,  db1 = W2^T dbc,  dW2 = dWc W1^T + dbc b1^T,  db2 = dbc.
//
// One persistent CTA per SM (4 warps; thread = doc) walks its slates of 128 docs = one 128-row MMA tile:
//   TMA      the slate's feature tile X[128, F] -> shared memory TWICE (5 boxes of 128 x 32 floats each): with the 128-byte
//            swizzle of 16-byte atoms for the forward MMA (X is its K-major A operand) and with the 128-byte swizzle of
//            32-byte atoms for the gradient MMA, which reads the same tile MN-major (docs are its K dimension) -- the only
//            layout the tensor core transposes 32-bit operands from.  The next slate's copies are in flight while this
//            one is worked on (the forward copy as soon as the forward MMA is done, the other after the gradient MMA);
//            the second copy of a tile comes out of L2: X is read from HBM ONCE.
//   MMA 1    tcgen05.mma kind::tf32, M = 128, N = 32, K = F: Z[128, 32] in TMEM  (A = X tile, B = Wc, both K-major)
//   epilogue tcgen05.ld: thread d owns row d of Z: sigmoid, dot with w3 -> the doc's score, a[32] stays in registers
//   loss     approxNDCG of the slate on the 128 scores (factorised pair sigmoid: one MUFU.RCP per ordered pair and
//            pass; direct form when the score range is too wide), gradient d loss / d s in the same registers
//   MMA 2    tcgen05.mma kind::tf32, M = 64, N = 160, K = 128 docs, ACCUMULATING over the CTA's slates in TMEM:
//            rows 0..31 = dZ^T (written to shared memory in the swizzled K-major layout, rows 32..63 stay zero);
//            B = the X tile read MN-major.  dbc = sum dZ and dw3 = sum a ds are sums of terms that cancel almost
//            completely (sum ds = 0 for approxNDCG), so they do NOT go through tf32: every thread accumulates its own
//            doc position's terms over the CTA's slates in fp32 registers, one shuffle reduction at the very end.
//   end      the CTA's accumulators -> partials in global memory; a second small kernel adds the partials in CTA order.
// Precision: tf32 operands (10-bit mantissa), fp32 accumulation: scores agree with the fp32 torch modules to ~1e-3
// relative (the tolerance the tests state), the loss arithmetic itself is the fp32 / MUFU arithmetic of K1.
#include <cuda.h>

#include "ltr_common.cuh"
#include "ltr_host.h"

namespace ltr {

constexpr int RF_DOCS = 128;           // docs per slate = rows of an MMA tile
constexpr int RF_H = 32;               // hidden units after the folded layers
constexpr int RF_CHUNK = 32;           // floats per 128-byte swizzle row
constexpr int RF_MAX_F = 160;          // features: at most 5 chunks of 32
constexpr int RF_KCHUNKS = 5;
constexpr int RF_X_CHUNK_BYTES = RF_DOCS * 128;                 // 16 KB: [128 docs][32 floats]
constexpr int RF_X_BYTES = RF_KCHUNKS * RF_X_CHUNK_BYTES;       // 80 KB per copy
constexpr int RF_W_CHUNK_BYTES = RF_H * 128;                    // 4 KB: [32 rows][32 floats]
constexpr int RF_W_BYTES = RF_KCHUNKS * RF_W_CHUNK_BYTES;       // 20 KB
constexpr int RF_DZ_ROWS = 64;                                  // rows of the second MMA's A operand
constexpr int RF_DZ_CHUNK_BYTES = RF_DZ_ROWS * 128;             // 8 KB: [64 rows][32 docs]
constexpr int RF_DZ_BYTES = 4 * RF_DZ_CHUNK_BYTES;              // 32 KB
constexpr int RF_N2 = 160;                                      // columns of the gradient accumulator: the 5 feature chunks (MN-major B: whole 32-float atoms)
constexpr int RF_TMEM_COLS = 256;                               // 32 (Z) + 144 (gradients) -> next power of two
constexpr int RF_THREADS = 128;
constexpr size_t RF_SMEM = 1024 + 2 * (size_t)RF_X_BYTES + RF_W_BYTES + RF_DZ_BYTES + 4096;

struct RankerFusedParams {
    const float* y_true;     // [B, 128]
    const float* bc;         // [32]
    const float* w3;         // [32]
    float b3;
    int B, F;
    float alpha, eps, pad;
    const double* disc_prefix;   // P[r] = sum_{q<=r} 1/log2(1+q), r = 0..128
    float* scores;           // [B, 128] out
    float* slate_ndcg;       // [B] out
    float* loss;             // [1] out
    float* part;             // [grid][64 * 144] per-CTA gradient accumulators
    float* part_misc;        // [grid][2]: sum of slate values, sum of ds
    float* part_bias;        // [grid][64]: dbc (32), dw3 (32) in fp32
    unsigned* counter;       // finished-CTA ticket (workspace, left at 0)
    int grid;
};

// ---- mbarrier / TMA / tcgen05 wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a barrier that never completes traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle: start address, leading / stride byte offsets (all >> 4), version 1
// (cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30), SBO [32,46), version [46,48), layout type [61,64) = 2)
// layout: 2 = SWIZZLE_128B (16-byte atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms; MN-major 32-bit operands)
__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint64_t umma_desc_sw128(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return umma_desc(smem_addr, lbo_bytes, sbo_bytes, 2u);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A / B tf32, majors, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr unsigned umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn_major << 15) | ((unsigned)b_mn_major << 16) |
           ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// byte offset of element (row, col) in a [rows][32 floats] tile with the 128-byte swizzle (tile base 1024-aligned)
__device__ __forceinline__ unsigned sw128_offset(int row, int col) {
    return (unsigned)row * 128u + ((((unsigned)col >> 2) ^ ((unsigned)row & 7u)) << 4) + (((unsigned)col & 3u) << 2);
}

struct RfShared {
    uint64_t fullk;          // TMA landed the forward copy of the slate's X tile
    uint64_t fullm;          // TMA landed the gradient-MMA copy
    uint64_t wfull;          // TMA landed Wc
    uint64_t mma1;           // the forward MMAs of the slate are complete (the forward copy is free)
    uint64_t mma2;           // the gradient MMAs of the slate are complete (the other copy and the dZ tile are free)
    unsigned tmem_base;
    int is_last;
    float s[RF_DOCS];        // compacted per-doc values of the loss: u, v, c
    float u[RF_DOCS];
    float v[RF_DOCS];
    float c[RF_DOCS];
    float red[8];
    float w3[RF_H], bc[RF_H];
    float fold[4][2 * RF_H];
    int wcount[4];
    int lvl[8];
};

__device__ __forceinline__ float rf_block_sum(float x, float* red) {
    x = warp_sum(x);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
}
__device__ __forceinline__ float rf_block_max(float x, float* red) {
    x = warp_max(x);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}

__global__ void __launch_bounds__(RF_THREADS, 1)
ranker_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xm,
                    const __grid_constant__ CUtensorMap map_w, const RankerFusedParams p) {
    extern __shared__ unsigned char rf_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(rf_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* Xk = base;                     // forward copy: 128-byte swizzle, 16-byte atoms
    unsigned char* Xm = base + RF_X_BYTES;        // gradient copy: 128-byte swizzle, 32-byte atoms
    unsigned char* wsm = base + 2 * RF_X_BYTES;
    unsigned char* dz = wsm + RF_W_BYTES;
    RfShared& sh = *reinterpret_cast<RfShared*>(dz + RF_DZ_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nslates = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // slates blockIdx.x, + grid, ...

    if (tid == 0) {
        mbar_init(&sh.fullk, 1);
        mbar_init(&sh.fullm, 1);
        mbar_init(&sh.wfull, 1);
        mbar_init(&sh.mma1, 1);
        mbar_init(&sh.mma2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                     "n"(RF_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = sh.tmem_base;
    const unsigned tmem_z = tmem, tmem_g = tmem + RF_H;          // columns [0, 32): Z; [32, 176): gradients

    // first tiles in flight: Wc (once) and the first slate
    if (tid == 0 && nslates > 0) {
        mbar_expect_tx(&sh.wfull, RF_W_BYTES);
        for (int ck = 0; ck < RF_KCHUNKS; ++ck) tma_load_2d(wsm + ck * RF_W_CHUNK_BYTES, &map_w, ck * RF_CHUNK, 0, &sh.wfull);
        mbar_expect_tx(&sh.fullk, RF_X_BYTES);
        for (int ck = 0; ck < RF_KCHUNKS; ++ck)
            tma_load_2d(Xk + ck * RF_X_CHUNK_BYTES, &map_x, ck * RF_CHUNK, (int)blockIdx.x * RF_DOCS, &sh.fullk);
        mbar_expect_tx(&sh.fullm, RF_X_BYTES);
        for (int ck = 0; ck < RF_KCHUNKS; ++ck)
            tma_load_2d(Xm + ck * RF_X_CHUNK_BYTES, &map_xm, ck * RF_CHUNK, (int)blockIdx.x * RF_DOCS, &sh.fullm);
    }
    if (tid < RF_H) sh.w3[tid] = __ldg(p.w3 + tid), sh.bc[tid] = __ldg(p.bc + tid);
    // rows 32..63 of the gradient MMA's A tile are never written again: zero once
    for (int i = tid; i < RF_DZ_BYTES / 16; i += RF_THREADS) reinterpret_cast<float4*>(dz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float gdz[RF_H], gaw[RF_H];                                  // fp32: this doc position's share of dbc and dw3
#pragma unroll
    for (int j = 0; j < RF_H; ++j) gdz[j] = 0.f, gaw[j] = 0.f;
    __syncthreads();
    const float scale = p.alpha * kLog2e;
    const bool clamp_pairs = (float)RF_DOCS * p.eps > 2e-7f;     // (below that the eps clamp of approxNDCG.py:49 is under fp32 resolution)
    const float invB = 1.0f / (float)p.B;
    const int ksteps1 = (p.F + 7) / 8;                           // k steps of 8 features
    constexpr unsigned IDESC1 = umma_idesc_tf32(128, RF_H, 0, 0);
    constexpr unsigned IDESC2 = umma_idesc_tf32(64, RF_N2, 0, 1);
    float val_sum = 0.f, ds_sum = 0.f;                           // thread 0: sum of slate values; all: own ds (reduced at the end)

    for (int it = 0; it < nslates; ++it) {
        const int slate = (int)blockIdx.x + it * (int)gridDim.x;
        const unsigned ph = (unsigned)it & 1u;
        const float y = __ldg(p.y_true + (size_t)slate * RF_DOCS + tid);
        // ---- forward MMA ----------------------------------------------------------------------------------------------
        if (it == 0) mbar_wait_bounded(&sh.wfull, 0);
        mbar_wait_bounded(&sh.fullk, ph);
        if (tid == 0) {
            tc_fence_after();
            const unsigned xa = smem_u32(Xk), wa = smem_u32(wsm);
            for (int ks = 0; ks < ksteps1; ++ks) {
                const unsigned off = (unsigned)(ks & 3) * 32u;
                tc_mma_tf32(tmem_z, umma_desc_sw128(xa + (ks >> 2) * RF_X_CHUNK_BYTES + off, 16, 1024),
                            umma_desc_sw128(wa + (ks >> 2) * RF_W_CHUNK_BYTES + off, 16, 1024), IDESC1, ks > 0);
            }
            tc_commit(&sh.mma1);
            // the next slate's gradient copy: its buffer was last read by the gradient MMAs of the previous iteration
            if (it >= 1) {
                mbar_wait_bounded(&sh.mma2, ph ^ 1u);
                mbar_expect_tx(&sh.fullm, RF_X_BYTES);
                for (int ck = 0; ck < RF_KCHUNKS; ++ck)
                    tma_load_2d(Xm + ck * RF_X_CHUNK_BYTES, &map_xm, ck * RF_CHUNK, slate * RF_DOCS, &sh.fullm);
            }
        }
        // ---- epilogue 1: row `tid` of Z -> a[32], score --------------------------------------------------------------
        mbar_wait_bounded(&sh.mma1, ph);
        tc_fence_after();
        if (tid == 0 && it + 1 < nslates) {                   // the forward copy is free: the next slate's goes in flight now
            mbar_expect_tx(&sh.fullk, RF_X_BYTES);
            for (int ck = 0; ck < RF_KCHUNKS; ++ck)
                tma_load_2d(Xk + ck * RF_X_CHUNK_BYTES, &map_x, ck * RF_CHUNK, (slate + (int)gridDim.x) * RF_DOCS, &sh.fullk);
        }
        float a[RF_H];
        tmem_ld32(tmem_z + ((unsigned)(warp * 32) << 16), a);
        float score = p.b3;
#pragma unroll
        for (int j = 0; j < RF_H; ++j) {
            const float z = a[j] + sh.bc[j];
            a[j] = 1.0f / (1.0f + __expf(-z));
            score = fmaf(a[j], sh.w3[j], score);
        }
        p.scores[(size_t)slate * RF_DOCS + tid] = score;
        // ---- approxNDCG of the slate (losses/approxNDCG.py:7-53) -------------------------------------------
        const bool valid = (y != p.pad);
        const unsigned vb = __ballot_sync(LTR_FULL_MASK, valid);
        const int lv = (int)fminf(fmaxf(y, -1.f), 5.f);
        const bool odd = valid && !((float)lv == y && lv >= 0 && lv <= 4);
        if (lane == 0) sh.wcount[warp] = __popc(vb);
        if (tid < 8) sh.lvl[tid] = 0;
        __syncthreads();
        if (odd) atomicOr(&sh.lvl[0], 1);
#pragma unroll
        for (int L = 1; L <= 4; ++L) {
            const unsigned bl = __ballot_sync(LTR_FULL_MASK, valid && lv == L);
            if (lane == 0 && bl) atomicAdd(&sh.lvl[L], __popc(bl));
        }
        __syncthreads();
        int before = __popc(vb & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) before += sh.wcount[w];
        const int m = sh.wcount[0] + sh.wcount[1] + sh.wcount[2] + sh.wcount[3];
        // ideal DCG over all positions: integer labels 0..4 -> level L fills the ranks [e_{L+1}, e_L)
        float maxdcg;
        if (!sh.lvl[0]) {
            const int e4 = sh.lvl[4], e3 = e4 + sh.lvl[3], e2 = e3 + sh.lvl[2], e1 = e2 + sh.lvl[1];
            const double P4 = p.disc_prefix[e4], P3 = p.disc_prefix[e3], P2 = p.disc_prefix[e2], P1 = p.disc_prefix[e1];
            maxdcg = (float)(15.0 * P4 + 7.0 * (P3 - P4) + 3.0 * (P2 - P3) + (P1 - P2));
        } else {
            // other labels: ideal order by a bitonic sort of the (clamped) label values
            sh.s[tid] = valid ? fmaxf(y, 0.f) : 0.f;
            __syncthreads();
            block_bitonic_desc(sh.s, RF_DOCS);
            const float term = (exp2f(sh.s[tid]) - 1.0f) / log2f(2.0f + (float)tid);
            maxdcg = rf_block_sum(term, sh.red);
            __syncthreads();
        }
        maxdcg = fmaxf(maxdcg, p.eps);
        const float G = valid ? (exp2f(fmaxf(y, 0.f)) - 1.0f) / maxdcg : 0.f;
        const float t = score * scale;
        const float ninf = __int_as_float(0xff800000), pinf = __int_as_float(0x7f800000);
        const float tmax = rf_block_max(valid ? t : ninf, sh.red);
        const float tmin = -rf_block_max(valid ? -t : ninf, sh.red);
        const bool fact = (tmax - tmin) < 120.0f;               // u v stays inside fp32; NaN / inf scores take the direct form
        const float mid = 0.5f * tmax + 0.5f * tmin;
        float myu = 0.f;
        if (valid) {
            myu = fact ? ex2_approx(t - mid) : t;
            sh.u[before] = myu;
            sh.v[before] = fact ? ex2_approx(mid - t) : t;
        }
        __syncthreads();
        float sum = 0.f;                                         // sum over ALL valid j (self included) of sigma_ij
        if (valid) {
            if (fact) {
                if (!clamp_pairs) {
#pragma unroll 4
                    for (int j = 0; j < m; ++j) sum += rcp_approx(fmaf(myu, sh.v[j], 1.0f));
                } else {
                    for (int j = 0; j < m; ++j) sum += fmaxf(rcp_approx(fmaf(myu, sh.v[j], 1.0f)), p.eps);
                }
            } else {
                for (int j = 0; j < m; ++j) {
                    const float r = rcp_approx(1.0f + ex2_capped(t - sh.v[j]));
                    sum += clamp_pairs ? fmaxf(r, p.eps) : r;
                }
            }
        }
        // pos_i = 1 + sum_{j != i}: the self term is sigma(0) = 1/2
        const float pos = 1.0f + sum - (clamp_pairs ? fmaxf(0.5f, p.eps) : 0.5f);
        const float lg = log2f(1.0f + pos);
        const float val = valid ? G / lg : 0.f;
        const float c = valid ? invB * G / (kLn2 * (1.0f + pos) * lg * lg) : 0.f;
        const float slate_val = rf_block_sum(val, sh.red);
        if (valid) sh.c[before] = c;
        __syncthreads();
        if (tid == 0) {
            p.slate_ndcg[slate] = slate_val;
            val_sum += slate_val;
        }
        float gacc = 0.f;                                        // sum_j w_ij (c_j - c_i)
        if (valid) {
            if (fact) {
                if (!clamp_pairs) {
#pragma unroll 4
                    for (int j = 0; j < m; ++j) {
                        const float pq = myu * sh.v[j];
                        const float r = rcp_approx(1.0f + pq);
                        gacc = fmaf(r * r * pq, sh.c[j] - c, gacc);
                    }
                } else {
                    for (int j = 0; j < m; ++j) {
                        const float pq = myu * sh.v[j];
                        const float r = rcp_approx(1.0f + pq);
                        gacc = fmaf((r > p.eps) ? r * r * pq : 0.f, sh.c[j] - c, gacc);
                    }
                }
            } else {
                for (int j = 0; j < m; ++j) {
                    const float pq = ex2_capped(t - sh.v[j]);
                    const float r = rcp_approx(1.0f + pq);
                    float w = r * (pq * r);
                    w = (pq == pinf) ? 0.f : w;
                    if (clamp_pairs && !(r > p.eps)) w = 0.f;
                    gacc = fmaf(w, sh.c[j] - c, gacc);
                }
            }
        }
        const float ds = valid ? p.alpha * gacc : 0.f;           // d loss / d score of this doc
        ds_sum += ds;
#ifdef RF_DEBUG_DS
        p.scores[(size_t)slate * RF_DOCS + tid] = ds;
#endif
        // ---- gradient MMA: rows 0..31 = dZ^T, rows 32..63 = (a ds)^T, K = the 128 docs of the slate ------------------------
        if (it >= 1) mbar_wait_bounded(&sh.mma2, ph ^ 1u);       // the dZ tile is free again
        mbar_wait_bounded(&sh.fullm, ph);                        // the gradient copy of this slate has landed
        {
            unsigned char* dzc = dz + (tid >> 5) * RF_DZ_CHUNK_BYTES;    // chunk of 32 docs
            const int col = tid & 31;
#pragma unroll
            for (int j = 0; j < RF_H; ++j) {
                const float aj = a[j];
                const float dzj = ds * sh.w3[j] * aj * (1.0f - aj);
                *reinterpret_cast<float*>(dzc + sw128_offset(j, col)) = dzj;
                gdz[j] += dzj;
                gaw[j] = fmaf(aj, ds, gaw[j]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA's reads
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const unsigned da = smem_u32(dz), xa = smem_u32(Xm);
            for (int ks = 0; ks < RF_DOCS / 8; ++ks) {
                // A: K-major, 8 docs = 32 bytes inside the 128-byte row of chunk ks / 4; B: the X tile MN-major: 8 docs = two
                // groups of 4 rows (512 bytes apart), feature chunks 16 KB apart (the leading-dimension offset)
                tc_mma_tf32(tmem_g, umma_desc_sw128(da + (ks >> 2) * RF_DZ_CHUNK_BYTES + (unsigned)(ks & 3) * 32u, 16, 1024),
                            umma_desc(xa + (unsigned)ks * 1024u, RF_X_CHUNK_BYTES, 512, 1u), IDESC2, (it > 0 || ks > 0));
            }
            tc_commit(&sh.mma2);
        }
    }
    // ---- the CTA's accumulators -> global partials --------------------------------------------------------------------
    if (nslates > 0) {
        mbar_wait_bounded(&sh.mma2, (unsigned)(nslates - 1) & 1u);
        tc_fence_after();
        // M = 64 accumulator: row r lives in lane (r % 16) + 32 (r / 16): lanes 0..15 of every warp's sub-partition
        float* mine = p.part + (size_t)blockIdx.x * (RF_DZ_ROWS * RF_N2);
        for (int c0 = 0; c0 < RF_N2; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_g + ((unsigned)(warp * 32) << 16) + (unsigned)c0, v);
            if (lane < 16) {
                const int row = warp * 16 + lane;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < RF_N2) mine[(size_t)row * RF_N2 + c0 + i] = v[i];
            }
        }
    } else {
        float* mine = p.part + (size_t)blockIdx.x * (RF_DZ_ROWS * RF_N2);
        for (int i = tid; i < RF_DZ_ROWS * RF_N2; i += RF_THREADS) mine[i] = 0.f;
    }
    const float ds_tot = rf_block_sum(ds_sum, sh.red);
#pragma unroll
    for (int j = 0; j < RF_H; ++j) {
        const float x = warp_sum(gdz[j]), w = warp_sum(gaw[j]);
        if (lane == 0) sh.fold[warp][j] = x, sh.fold[warp][RF_H + j] = w;
    }
    __syncthreads();
    if (tid < 2 * RF_H)
        p.part_bias[(size_t)blockIdx.x * (2 * RF_H) + tid] = (sh.fold[0][tid] + sh.fold[1][tid]) + (sh.fold[2][tid] + sh.fold[3][tid]);
    if (tid == 0) {
        p.part_misc[2 * blockIdx.x + 0] = val_sum;
        p.part_misc[2 * blockIdx.x + 1] = ds_tot;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(RF_TMEM_COLS) : "memory");
}

// partials of the CTAs, added in CTA order: gWc [32, F], gbc [32], gw3 [32], gb3 [1], loss [1]
__global__ void __launch_bounds__(256) ranker_fused_reduce_kernel(const RankerFusedParams p, float* gWc, float* gbc, float* gw3,
                                                                  float* gb3) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int total = RF_DZ_ROWS * RF_N2;
    if (idx < total) {
        float t = 0.f;
        for (int g = 0; g < p.grid; ++g) t += p.part[(size_t)g * total + idx];
        const int row = idx / RF_N2, col = idx % RF_N2;
        if (row < RF_H && col < p.F) gWc[row * p.F + col] = t;
    }
    if (idx < 2 * RF_H) {
        float t = 0.f;
        for (int g = 0; g < p.grid; ++g) t += p.part_bias[(size_t)g * (2 * RF_H) + idx];
        if (idx < RF_H) gbc[idx] = t;
        else gw3[idx - RF_H] = t;
    }
    if (idx == 0) {
        double v = 0.0, d = 0.0;
        for (int g = 0; g < p.grid; ++g) v += (double)p.part_misc[2 * g], d += (double)p.part_misc[2 * g + 1];
        p.loss[0] = (float)(-v / (double)p.B);
        gb3[0] = (float)d;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [rows, F] fp32 row-major -> boxes of box_rows x 32 floats, 128-byte swizzle, zero fill beyond the tensor
static int make_map(CUtensorMap* map, const float* ptr, unsigned long long rows, int F, int box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return set_error(LTR_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)F * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)RF_CHUNK, (cuuint32_t)box_rows};
    cuuint32_t elem[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, elem,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(LTR_EINVAL, "cuTensorMapEncodeTiled failed (alignment / shape of the feature tensor)");
    return 0;
}

}  // namespace ltr

LTR_API size_t ltr_triple_scratch_bytes(int B, int n, int F) {
    (void)n;
    (void)F;
    if (B <= 0) return 0;
    const size_t grid = (size_t)ltr::device_sm_count();
    return grid * (ltr::RF_DZ_ROWS * ltr::RF_N2 + 2 + 2 * ltr::RF_H) * sizeof(float);
}

LTR_API int ltr_triple_approx_ndcg_fwdbwd(const float* X, const float* y_true, const float* Wc, const float* bc, const float* w3,
                                          float b3, int B, int n, int F, float alpha, float eps, float pad_value,
                                          const double* disc_prefix, int disc_len, float* scores, float* slate_ndcg, float* loss,
                                          float* gWc, float* gbc, float* gw3, float* gb3, void* scratch, size_t scratch_bytes,
                                          void* workspace, void* stream) {
    using namespace ltr;
    if (!X || !y_true || !Wc || !bc || !w3 || !scores || !slate_ndcg || !loss || !gWc || !gbc || !gw3 || !gb3 || !scratch ||
        !workspace || !disc_prefix || B <= 0)
        return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: null pointer or empty batch");
    if (n != RF_DOCS) return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: slates of exactly 128 docs (one MMA tile)");
    if (F < 4 || F > RF_MAX_F || F % 4)
        return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: 4 <= F <= 160, F a multiple of 4 (16-byte rows for TMA)");
    if (disc_len < RF_DOCS + 1) return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: discount table shorter than 129");
    if (!(alpha > 0.f)) return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: alpha must be positive");
    if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wc)) % 16)
        return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: X and Wc must be 16-byte aligned");
    if (scratch_bytes < ltr_triple_scratch_bytes(B, n, F)) return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: scratch too small");
    int grid = device_sm_count();
    if (grid > B) grid = B;
    CUtensorMap map_x, map_xm, map_w;
    if (int rc = make_map(&map_x, X, (unsigned long long)B * RF_DOCS, F, RF_DOCS, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map(&map_xm, X, (unsigned long long)B * RF_DOCS, F, RF_DOCS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_map(&map_w, Wc, RF_H, F, RF_H, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    RankerFusedParams p;
    p.y_true = y_true; p.bc = bc; p.w3 = w3; p.b3 = b3; p.B = B; p.F = F;
    p.alpha = alpha; p.eps = eps; p.pad = pad_value; p.disc_prefix = disc_prefix;
    p.scores = scores; p.slate_ndcg = slate_ndcg; p.loss = loss;
    p.part = reinterpret_cast<float*>(scratch);
    p.part_misc = p.part + (size_t)device_sm_count() * (RF_DZ_ROWS * RF_N2);
    p.part_bias = p.part_misc + (size_t)device_sm_count() * 2;
    p.counter = nullptr;
    p.grid = grid;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    LTR_CUDA(cudaFuncSetAttribute(ranker_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RF_SMEM));
    ranker_fused_kernel<<<grid, RF_THREADS, RF_SMEM, st>>>(map_x, map_xm, map_w, p);
    int rc = (int)cudaGetLastError();
    if (rc) return set_cuda_error(rc, "ltr_triple_approx_ndcg_fwdbwd");
    ranker_fused_reduce_kernel<<<(RF_DZ_ROWS * RF_N2 + 255) / 256, 256, 0, st>>>(p, gWc, gbc, gw3, gb3);
    rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_triple_approx_ndcg_fwdbwd (reduce)") : 0;
}
