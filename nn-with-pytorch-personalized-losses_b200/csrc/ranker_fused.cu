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
// One persistent, warp-specialised CTA per SM; a slate of 128 docs = one 128-row MMA tile.  The CTA's slates are dealt
// round robin to G compute groups of 4 warps (thread = doc, warp = one TMEM lane quarter), so G slates are in different
// phases at any time; three warp-uniform roles (one elected lane issues) feed them:
//   loader      TMA through two rings.  F ring: chunks of [128 docs][32 features] = 16 KB with the 128-byte swizzle of
//               16-byte atoms: X is the K-major A operand of the forward MMA; these are the HBM reads (L2 evict_last).
//               D ring: the same bytes again, out of L2 (evict_first), as slots of [32 docs][all feature chunks] = 5 x 4 KB
//               with the 128-byte swizzle of 32-byte atoms: the gradient MMA reads them MN-major -- docs are its K
//               dimension -- and that is the only layout the tensor core transposes 32-bit operands from.  X is read from
//               HBM ONCE (ncu: 290 MB for 285 MB of features).
//   F issuer    tcgen05.mma kind::tf32, M = 128, N = 32, K = F: Z[128, 32] of the slate into its group's TMEM columns, when
//               the group is through with the loss of its previous slate; every chunk's slot is released by tcgen05.commit.
//   D issuer    tcgen05.mma kind::tf32, M = 64, N = 32 x chunks (160), K = 32 docs per slot, ACCUMULATING over all of the
//               CTA's slates in TMEM (in slate order: bitwise reproducible): A = dZ^T of the slate (rows 0..31, K-major,
//               written by the group; rows 32..63 of the M = 64 tile alias whatever follows and land in accumulator rows
//               nobody reads), B = the X slot read MN-major.  One MMA spans all features, so A is read once per k step.
//   group       tcgen05.ld: thread d owns row d of Z: sigmoid, dot with w3 -> the doc's score (a[32] stays in registers);
//               approxNDCG of the slate exactly as the K1 kernel computes it (approx_pair_tiles.cuh: factorised pair
//               sigmoid, every unordered pair once per pass, block-circulant pair schedule over the group's 4 warps;
//               direct form when the score range is too wide or the eps clamp matters); dZ of the doc -> shared memory
//               for the D issuer.  dbc = sum dZ and dw3 = sum a ds are sums of terms that cancel almost completely
//               (sum ds = 0 for approxNDCG), so they do NOT go through tf32: a butterfly of warp shuffles leaves lane j
//               with the warp's sum for hidden unit j, accumulated in fp32 over the group's slates.
//   end         the CTA's accumulators -> partials in global memory; a second small kernel adds the partials in CTA order.
// The barrier protocol (mbarrier parity waits are only sound within one phase) is modelled in tests/test_ranker_protocol_cpu.py.
// Bound: every feature byte crosses L2 -> SM twice (570 MB per 4096 x 128 x 136 batch; `ltr_probe_tma_stream` measures
// 6.4 TB/s for the HBM pass and 11.5 TB/s for an L2-resident pass on this part), see DESIGN.md.
// Precision: tf32 operands (10-bit mantissa), fp32 accumulation: scores agree with the fp32 torch modules to ~1e-3
// relative (the tolerance the tests state), the loss arithmetic itself is the fp32 / MUFU arithmetic of K1.
#include <cuda.h>

#include "approx_pair_tiles.cuh"
#include "ltr_common.cuh"
#include "ltr_host.h"

#ifndef RF_GROUPS
#define RF_GROUPS 3          // compute groups per CTA (A/B knob; measured 4096 x 128 x 136: 2 groups 85.3 us, 3: 84.6 us, 4: 86.5 us)
#endif

namespace ltr {

constexpr int RF_DOCS = 128;           // docs per slate = rows of an MMA tile
constexpr int RF_H = 32;               // hidden units after the folded layers
constexpr int RF_CHUNK = 32;           // floats per 128-byte swizzle row
constexpr int RF_MAX_F = 160;          // features: at most 5 chunks of 32
constexpr int RF_KCHUNKS = 5;
constexpr int RF_SLOT_BYTES = RF_DOCS * 128;                    // 16 KB: [128 docs][32 floats]
constexpr int RF_W_CHUNK_BYTES = RF_H * 128;                    // 4 KB: [32 rows][32 floats]
constexpr int RF_W_BYTES = RF_KCHUNKS * RF_W_CHUNK_BYTES;       // 20 KB
constexpr int RF_DZ_CHUNK_BYTES = RF_H * 128;                   // 4 KB: [32 hidden units][32 docs]
constexpr int RF_DZ_TILE_BYTES = 4 * RF_DZ_CHUNK_BYTES;         // 16 KB: dZ^T of one slate
// dZ tiles, used in slate order by whichever group owns the slate
constexpr int RF_NDZ = (RF_GROUPS <= 3) ? 3 : 2;
static_assert(RF_GROUPS <= 2 * RF_NDZ, "the dZ tile hand-over needs G <= 2 RF_NDZ");
constexpr int RF_N2 = RF_KCHUNKS * RF_CHUNK;                    // columns of the gradient accumulator
constexpr int RF_G = RF_GROUPS;
#ifndef RF_NSF_SLOTS
#define RF_NSF_SLOTS 5
#endif
#ifndef RF_NSD_SLOTS
#define RF_NSD_SLOTS 3
#endif
constexpr int RF_NSF = RF_NSF_SLOTS;                            // F ring: forward chunks in flight (HBM latency)
constexpr int RF_QDOCS = 32;                                    // D ring: a slot = 32 docs x ALL feature chunks ...
constexpr int RF_QBOX_BYTES = RF_QDOCS * 128;                   // ... 4 KB per chunk: [32 docs][32 floats] ...
constexpr int RF_QSLOT_BYTES = RF_KCHUNKS * RF_QBOX_BYTES;      // ... 20 KB per slot, so that one MMA spans all N = 160 features
constexpr int RF_NSD = RF_NSD_SLOTS;                            // D ring of doc quarters (L2 hits); shared memory decides
#ifndef RF_WAIT_HINT_NS
#define RF_WAIT_HINT_NS 4000
#endif
#ifndef RF_LOADER_D_FIRST
#define RF_LOADER_D_FIRST 0
#endif
#ifndef RF_EARLY_F
#define RF_EARLY_F 0          // 1: the group's next forward MMA goes as soon as Z is in registers (measured: same time, 25 % of
                             // the gradient job's second reads then miss L2)
#endif
constexpr int RF_THREADS = RF_G * 128 + 96;                     // compute groups + loader, F issuer, D issuer warps
constexpr int RF_TMEM_COLS = (RF_G * RF_H + RF_N2 <= 256) ? 256 : 512;
constexpr int RF_PART = RF_H * RF_N2;                           // floats of one CTA's gradient partial

struct RankerFusedParams {
    const float* y_true;     // [B, 128]
    const float* bc;         // [32]
    const float* w3;         // [32]
    float b3;                // used when b3_ptr is null
    const float* b3_ptr;     // [1] (the whole-step entry: the bias stays on the device)
    int B, F;
    float alpha, eps, pad;
    const double* disc_prefix;   // P[r] = sum_{q<=r} 1/log2(1+q), r = 0..128
    float* scores;           // [B, 128] out
    float* slate_ndcg;       // [B] out
    float* loss;             // [1] out
    float* part;             // [grid][32 * 160] per-CTA gradient accumulators
    float* part_misc;        // [grid][2]: sum of slate values, sum of ds
    float* part_bias;        // [grid][64]: dbc (32), dw3 (32) in fp32
    int grid;
    int loss_kind;           // 0: approxNDCG (loss = - mean of the slate values), 1: ListNet (loss = - sum)
    unsigned long long* trace;   // RF_TRACE builds: [role][event][slate] globaltimer stamps of CTA 0 (diagnostics)
};

#ifdef RF_TRACE
#define RF_TRACE_MAX_K 64
#define RF_STAMP(role, ev, k)                                                                                  \
    do {                                                                                                      \
        if (p.trace && blockIdx.x == 0 && (k) < RF_TRACE_MAX_K) {                                            \
            unsigned long long t_;                                                                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                           \
            p.trace[((role) * 8 + (ev)) * RF_TRACE_MAX_K + (k)] = t_;                                         \
        }                                                                                                     \
    } while (0)
#else
#define RF_STAMP(role, ev, k) do {} while (0)
#endif

// ---- mbarrier / TMA / tcgen05 wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a barrier that never completes traps instead of hanging the GPU.  The suspend-time hint lets the
// hardware park the warp (up to ~4 us per try) instead of spinning through issue slots the compute groups need.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)RF_WAIT_HINT_NS)
            : "memory");
        if (spin > (1u << 21)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// with an L2 eviction-priority hint (createpolicy)
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
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

// one lane of a converged warp (warp-uniform code around it keeps the MMA operands in uniform registers: a branch on
// lane == 0 makes the compiler wrap every tcgen05 instruction in a ~14-instruction R2UR / vote loop)
__device__ __forceinline__ bool elect_one() {
    unsigned pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// per compute group: the slate state of the loss (the layout of the K1 kernel's SymShared)
struct __align__(16) RfGroup {
    float u[RF_DOCS];             // 2^(t - c)   (direct form: raw score)
    float v[RF_DOCS];             // 2^-(t - c)
    float c[RF_DOCS];             // d loss / d pos_i
    float part[4][RF_DOCS];       // per-warp partial sums of every doc (ranks, then gradients)
    int4 wstat[4];                // per warp: {valid count | bad-label flag << 16, min key, max key, packed label counts}
    float red[16];                // 4 slots of 4 per-warp values (the ListNet block runs several reductions per slate)
};

struct RfCtl {
    uint64_t full_f[RF_NSF], empty_f[RF_NSF];     // F ring: TMA landed / the chunk's MMAs are complete
    uint64_t full_d[RF_NSD], empty_d[RF_NSD];     // D ring
    uint64_t wfull;                               // TMA landed Wc
    uint64_t zfull[RF_G];                         // the forward MMAs of the group's slate are complete
    uint64_t zfree[RF_G];                         // the group is done with its slate: its next forward MMA may go (128 arrivals)
    uint64_t dzfull[RF_NDZ];                      // dZ of a slate has been written to the tile (128 arrivals)
    uint64_t dzfree[RF_NDZ];                      // the gradient MMAs reading the tile are complete
    uint64_t dzmine[RF_G];                        // the gradient MMAs of the group's latest slate are complete
    uint64_t alldone;                             // every MMA of the CTA is complete
    unsigned tmem_base;
    float w3[RF_H], bc[RF_H];
    float disc[RF_DOCS + 1];                      // discount prefix sums, rounded once to fp32 (as K1's constant table)
};

// sum over the warp's lanes of v[j], left in lane j: 31 exchanges (16 + 8 + 4 + 2 + 1); v is destroyed
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool hi = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = hi ? v[i] : v[i + s];
            const float keep = hi ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(LTR_FULL_MASK, send, s);
        }
    }
    return v[0];
}

// soft rank -> slate value and d loss / d pos (sh.c); one group barrier.  `sum` = sum_{j != i} sigma_ij of this thread's doc.
__device__ __forceinline__ float rf_middle(RfGroup& sh, int bar, int lane, int warp, bool valid, int dst, float gain, float coef,
                                           float sum) {
    float ndcg_part = 0.f;
    if (valid) {
        const float pos1 = 2.0f + sum;                       // 1 + pos_i
        const float l2 = lg2_approx(pos1);
        const float il2 = rcp_approx(l2);
        ndcg_part = gain * il2;
        sh.c[dst] = coef * gain * il2 * il2 * rcp_approx(pos1);
    }
    ndcg_part = warp_sum(ndcg_part);
    if (lane == 0) sh.red[warp] = ndcg_part;
    named_bar_sync(bar, 128);
    return (sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]);
}

// One slate in the factorised form, R docs per lane (the schedule of sym_slate in approx_ndcg_sym.cuh, on a group of 4 warps
// with a named barrier).  Returns sum_j w_ij (c_j - c_i) of this thread's doc; slate_val = the slate's approxNDCG.
template <int R>
__device__ __forceinline__ float rf_sym_slate(RfGroup& sh, int bar, int gt, int rot, int m, bool valid, int dst, float gain,
                                              float coef, float& slate_val) {
    const int lane = gt & 31, warp = gt >> 5;
    const int T = (m + R - 1) / R;            // active lanes = row blocks
    const int S = T >> 1;                     // circulant steps D = 1 .. T/2 (D = 0 is the in-block triangle)
    const int role = (warp + rot) & 3;
    const int d_lo = 1 + ((role * S) >> 2);
    const int d_hi = 1 + (((role + 1) * S) >> 2);
    const int steps = d_hi - d_lo;
    // for even T the antipodal step D = T/2 meets every pair twice and is weighted 1/2: it is the last step of role 3
    const bool anti = ((T & 1) == 0) && (role == 3) && (steps > 0);
    const int full = steps - (anti ? 1 : 0);
    const bool active = lane < T;
    const int src = active ? ((lane + 1 == T) ? 0 : lane + 1) : lane;
    int cb = lane + d_lo;
    if (cb >= T) cb -= T;
    if (!active) cb = 0;
    int ret = lane - (d_hi - 1);              // the lane that ends up holding this lane's column accumulators
    if (ret < 0) ret += T;
    if (steps == 0 || !active) ret = lane;    // nothing accumulated: the column sums are zero

    float urow[R], vcol[R], acc_r[R], acc_r_hi[R], acc_c[R];
    lds_vec<R>(sh.u + lane * R, urow);
    lds_vec<R>(sh.v + cb * R, vcol);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc_r[r] = 0.f;
        acc_r_hi[r] = 0.f;
        acc_c[r] = 0.f;
    }
    // ---------------- rank pass: one MUFU.RCP per unordered pair -----------------------------
    if (role == 0) {                          // pairs inside the lane's own block
        float vrow[R];
        lds_vec<R>(sh.v + lane * R, vrow);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = r + 1; c < R; ++c) {
                const float pp = urow[r] * vrow[c];
                const float rr = rcp_approx(pp + 1.0f);
                acc_r[r] += rr;
                acc_r[c] = fmaf(pp, rr, acc_r[c]);
            }
        }
    }
    // ONE copy of the tile code per pass (the antipodal step is the loop's last iteration with weight 1/2): the kernel's
    // instruction footprint, not its arithmetic, is what the SMs run out of first
#pragma unroll 1
    for (int i = 0; i < steps; ++i) {
        if (i) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
            }
        }
        rank_tile<R>(urow, vcol, (anti && i == full) ? 0.5f : 1.0f, acc_r, acc_r_hi, acc_c);
    }
    // the column accumulators go home; a doc's partial is (sum sigma_ij as a row) - (sum sigma_ij as a column)
    {
        float out[R];
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = (acc_r[r] + acc_r_hi[r]) - __shfl_sync(LTR_FULL_MASK, acc_c[r], ret);
        if (active) sts_vec<R>(sh.part[warp] + lane * R, out);
    }
    named_bar_sync(bar, 128);

    // ---------------- per doc: soft rank, slate value, d loss / d pos --------------------------
    float sum = 0.f;
    if (valid) {
        const float parts = (sh.part[0][dst] + sh.part[1][dst]) + (sh.part[2][dst] + sh.part[3][dst]);
        // column side: sum sigma_ji = (#rows met, antipodal ones at 1/2) - sum sigma_ij = R (T-1)/2 - cols.
        // padded tail of the last block (u = v = 0): as a column it reads sigma_ij = 1 on the row side.
        const int npadc = T * R - m;
        const int e = T - 1 - dst / R;                   // circulant distance to the last block
        const float f = (2 * e == T) ? 0.5f : ((2 * e < T) ? 1.0f : 0.0f);
        sum = (parts - f * (float)npadc) + 0.5f * (float)(R * (T - 1));
    }
    slate_val = rf_middle(sh, bar, lane, warp, valid, dst, gain, coef, sum);

    // ---------------- gradient pass: sigma recomputed, one MUFU.RCP per unordered pair ---------
    float crow[R], ccol[R];
    lds_vec<R>(sh.c + lane * R, crow);
    lds_vec<R>(sh.c + cb * R, ccol);
    lds_vec<R>(sh.v + cb * R, vcol);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc_r[r] = 0.f;
        acc_r_hi[r] = 0.f;
        acc_c[r] = 0.f;
    }
    if (role == 0) {
        float vrow[R];
        lds_vec<R>(sh.v + lane * R, vrow);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = r + 1; c < R; ++c) {
                const float pp = urow[r] * vrow[c];
                const float rr = rcp_approx(pp + 1.0f);
                const float e = (rr * (pp * rr)) * (crow[c] - crow[r]);
                acc_r[r] += e;
                acc_r[c] -= e;
            }
        }
    }
#pragma unroll 1
    for (int i = 0; i < steps; ++i) {
        if (i) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                vcol[r] = __shfl_sync(LTR_FULL_MASK, vcol[r], src);
                ccol[r] = __shfl_sync(LTR_FULL_MASK, ccol[r], src);
                acc_c[r] = __shfl_sync(LTR_FULL_MASK, acc_c[r], src);
            }
        }
        if (anti && i == full) {                // the antipodal step: weight 1/2 goes onto the c differences (last step)
#pragma unroll
            for (int r = 0; r < R; ++r) crow[r] *= 0.5f, ccol[r] *= 0.5f;
        }
        grad_tile<R>(urow, vcol, crow, ccol, acc_r, acc_r_hi, acc_c);
    }
    {
        float out[R];
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = (acc_r[r] + acc_r_hi[r]) + __shfl_sync(LTR_FULL_MASK, acc_c[r], ret);
        if (active) sts_vec<R>(sh.part[warp] + lane * R, out);
    }
    named_bar_sync(bar, 128);
    float sum2 = 0.f;
    if (valid) sum2 = (sh.part[0][dst] + sh.part[1][dst]) + (sh.part[2][dst] + sh.part[3][dst]);
    return sum2;
}

// Slates whose scaled score range does not fit the factorised form, or whose eps clamp is above fp32 resolution (both
// rare): one doc per thread, p = ex2(min(alpha log2e (s_i - s_j), 126)) per ordered pair, recomputed in the gradient
// pass.  sh.u holds the raw scores of the m compacted docs.  Out of line: keeps the common path's code small.
__device__ __noinline__ float rf_direct_rank_sum(const float* raw, int m, float scale, float si, float eps, int clamp) {
    float sum = 0.f;
    for (int j = 0; j < m; ++j) {
        const float rr = rcp_approx(ex2_capped((si - raw[j]) * scale) + 1.0f);
        sum += clamp ? fmaxf(rr, eps) : rr;
    }
    return sum - (clamp ? fmaxf(0.5f, eps) : 0.5f);             // the j == i term is exactly 1/2 (approxNDCG.py:34 drops it)
}
__device__ __noinline__ float rf_direct_grad_sum(const float* raw, const float* c, int m, float scale, float si, float ci,
                                                 float eps, int clamp) {
    const float pinf = __int_as_float(0x7f800000);
    float acc = 0.f;
    for (int j = 0; j < m; ++j) {
        const float pp = ex2_capped((si - raw[j]) * scale);
        const float rr = rcp_approx(pp + 1.0f);
        float w = rr * (pp * rr);
        w = (pp == pinf) ? 0.f : w;
        if (clamp && !(rr > eps)) w = 0.f;                       // the clamp is active: no gradient through this pair
        acc = fmaf(w, c[j] - ci, acc);
    }
    return acc;
}

template <int G>
__global__ void __launch_bounds__(G * 128 + 96, 1)
ranker_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xm,
                    const __grid_constant__ CUtensorMap map_w, const RankerFusedParams p) {
    extern __shared__ unsigned char rf_smem_raw[];
    // (aligned by an offset on the __shared__ array itself: a round trip through uintptr_t makes every access a generic LD / ST)
    unsigned char* base = rf_smem_raw + ((1024u - (smem_u32(rf_smem_raw) & 1023u)) & 1023u);
    unsigned char* fring = base;                                  // forward copies: 128-byte swizzle, 16-byte atoms
    unsigned char* dring = fring + RF_NSF * RF_SLOT_BYTES;        // gradient copies: 128-byte swizzle, 32-byte atoms
    unsigned char* wsm = dring + RF_NSD * RF_QSLOT_BYTES;
    unsigned char* dz = wsm + RF_W_BYTES;
    RfGroup* groups = reinterpret_cast<RfGroup*>(dz + RF_NDZ * RF_DZ_TILE_BYTES);
    RfCtl& ctl = *reinterpret_cast<RfCtl*>(groups + G);
    const int tid = threadIdx.x, lane = tid & 31, warp_id = tid >> 5;
    const int nsl = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;          // slates blockIdx.x, + grid, ...
    const int nch = (p.F + RF_CHUNK - 1) / RF_CHUNK;              // feature chunks
    const int ksteps = (p.F + 7) / 8;                             // forward k steps of 8 features

    if (tid == 0) {
        for (int i = 0; i < RF_NSF; ++i) mbar_init(&ctl.full_f[i], 1), mbar_init(&ctl.empty_f[i], 1);
        for (int i = 0; i < RF_NSD; ++i) mbar_init(&ctl.full_d[i], 1), mbar_init(&ctl.empty_d[i], 1);
        mbar_init(&ctl.wfull, 1);
        for (int g = 0; g < G; ++g) mbar_init(&ctl.zfull[g], 1), mbar_init(&ctl.zfree[g], 128), mbar_init(&ctl.dzmine[g], 1);
        for (int t = 0; t < RF_NDZ; ++t) mbar_init(&ctl.dzfull[t], 128), mbar_init(&ctl.dzfree[t], 1);
        mbar_init(&ctl.alldone, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_id == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)),
                     "n"(RF_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("griddepcontrol.launch_dependents;");           // the reduce kernel's blocks may become resident (they wait)
    asm volatile("griddepcontrol.wait;" ::: "memory");           // (whole-step entry: Wc / bc come from the fold kernel before us)
    if (tid < RF_H) ctl.w3[tid] = __ldg(p.w3 + tid), ctl.bc[tid] = __ldg(p.bc + tid);
    for (int i = tid; i <= RF_DOCS; i += blockDim.x) ctl.disc[i] = (float)p.disc_prefix[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = ctl.tmem_base;
    const unsigned tmem_g = tmem + G * RF_H;                      // columns [0, 32 G): Z of the groups; then the gradients

    if (warp_id < 4 * G) {
        // ================= compute group: thread = doc =====================================================================
        const int g = warp_id >> 2, w = warp_id & 3, gt = tid & 127;
        RfGroup& sh = groups[g];
        const int bar = 1 + g;
        unsigned dzoff[8];                                        // swizzled byte offset of column `lane` in a row of class j & 7
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) dzoff[jj] = ((((unsigned)lane >> 2) ^ (unsigned)jj) << 4) + (((unsigned)lane & 3u) << 2);
        const float scale = p.alpha * kLog2e;
        const int clamp_pairs = ((float)RF_DOCS * p.eps > 2e-7f) ? 1 : 0;   // (below that the eps clamp of approxNDCG.py:49 is under fp32 resolution)
        const float coef = (1.0f / (float)p.B) * (1.0f / kLn2);
        float acc_bc = 0.f, acc_w3 = 0.f;                         // lane j: this warp's share of dbc[j], dw3[j]
        float ds_sum = 0.f, val_sum = 0.f;
        const float b3 = p.b3_ptr ? __ldg(p.b3_ptr) : p.b3;
        float ynext = p.pad;
        if (g < nsl) ynext = ld_stream(p.y_true + (size_t)((int)blockIdx.x + g * (int)gridDim.x) * RF_DOCS + gt);
        int it = 0;
        for (int k = g; k < nsl; k += G, ++it) {
            const int slate = (int)blockIdx.x + k * (int)gridDim.x;
            const float yv = ynext;
            if (k + G < nsl) ynext = ld_stream(p.y_true + (size_t)(slate + G * (int)gridDim.x) * RF_DOCS + gt);
            // ---- epilogue of the forward MMA: row `gt` of Z -> a[32], score ---------------------------------------------
            if (gt == 0) RF_STAMP(g, 0, k);
            mbar_wait_bounded(&ctl.zfull[g], (unsigned)it & 1u);
            if (gt == 0) RF_STAMP(g, 1, k);
            tc_fence_after();
            float a[RF_H];
            tmem_ld32(tmem + (unsigned)(g * RF_H) + ((unsigned)(w * 32) << 16), a);
            tc_fence_before();
#if RF_EARLY_F
            // Z is in registers: the group's next forward MMA may go.  The F ring then drains as fast as HBM delivers instead of
            // at the pace of the groups; the price is the distance in L2 between a chunk's first read and its second (by the
            // gradient job, (2 G + 1) slates x 148 SMs later), which the eviction hints of the loader have to carry.
            mbar_arrive(&ctl.zfree[g]);
#endif
            float sv = b3;
#pragma unroll
            for (int j = 0; j < RF_H; ++j) {
                const float z = a[j] + ctl.bc[j];
                a[j] = rcp_approx(1.0f + ex2_approx(-kLog2e * z));
                sv = fmaf(a[j], ctl.w3[j], sv);
            }
            p.scores[(size_t)slate * RF_DOCS + gt] = sv;
            if (gt == 0) RF_STAMP(g, 2, k);
            float ds = 0.f, slate_val = 0.f;
            if (p.loss_kind == 1) {
                // ---- ListNet of the slate (losses/listnet.py:5-16): - sum_i softmax(y_true)_i log softmax(s)_i over ALL 128 docs --
                // (the reference gives padding no special treatment: a label of -1 is a label);  d loss / d s = softmax(s) - softmax(y)
                float* red = sh.red;                                 // 4 x 4 slots (red + pad_): one per reduction of the slate
                float x = warp_max(sv), yx = warp_max(yv);
                if (lane == 0) red[w] = x, red[4 + w] = yx;
                named_bar_sync(bar, 128);
                const float smax = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
                const float ymax = fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7]));
                const float es = ex2_approx((sv - smax) * kLog2e), ey = ex2_approx((yv - ymax) * kLog2e);
                x = warp_sum(es), yx = warp_sum(ey);
                if (lane == 0) red[8 + w] = x, red[12 + w] = yx;
                named_bar_sync(bar, 128);
                const float ssum = (red[8] + red[9]) + (red[10] + red[11]), ysum = (red[12] + red[13]) + (red[14] + red[15]);
                const float pt = __fdividef(ey, ysum);
                const float logp = (sv - smax) - lg2_approx(ssum) * kLn2;
                x = warp_sum(pt * logp);
                named_bar_sync(bar, 128);                            // (everyone has read slots 0..7)
                if (lane == 0) red[w] = x;
                named_bar_sync(bar, 128);
                slate_val = (red[0] + red[1]) + (red[2] + red[3]);
                ds = __fdividef(es, ssum) - pt;
                named_bar_sync(bar, 128);                            // (the next slate's first reduction writes slot 0 again)
            } else {
                // ---- approxNDCG of the slate (losses/approxNDCG.py:7-53; the K1 kernel's steps) --------------------------------
                const bool valid = (yv != p.pad);
                const float yc = fmaxf(yv, 0.f);                      // approxNDCG.py:37-38 clamp
                const int lvl = (int)fminf(yc, 100.f);
                const bool small_int = (yc == (float)lvl) && (lvl <= 4);
                const unsigned bal = __ballot_sync(LTR_FULL_MASK, valid);
                const unsigned bad = __ballot_sync(LTR_FULL_MASK, valid && !small_int);
                const int key = float_order_key(sv);
                const int kmin = __reduce_min_sync(LTR_FULL_MASK, valid ? key : 0x7fffffff);
                const int kmax = __reduce_max_sync(LTR_FULL_MASK, valid ? key : (int)0x80000000);
                const unsigned field = (valid && small_int && lvl >= 1) ? (1u << (8 * (lvl - 1))) : 0u;
                const unsigned cnts = __reduce_add_sync(LTR_FULL_MASK, field);
                if (lane == 0) sh.wstat[w] = make_int4(__popc(bal) | (bad ? 0x10000 : 0), kmin, kmax, (int)cnts);
                named_bar_sync(bar, 128);
                int m = 0, off = 0, any_bad = 0;
                int gmin = 0x7fffffff, gmax = (int)0x80000000;
                unsigned packed = 0;
    #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int4 st = sh.wstat[q];
                    const int cnt = st.x & 0xffff;
                    off += (q < w) ? cnt : 0;
                    m += cnt;
                    any_bad |= st.x >> 16;
                    gmin = min(gmin, st.y);
                    gmax = max(gmax, st.z);
                    packed += (unsigned)st.w;
                }
                const int dst = off + __popc(bal & ((1u << lane) - 1u));
                if (m > 0) {
                    // ideal DCG over ALL positions (approxNDCG.py:41-43)
                    float maxdcg, gain;
                    if (!any_bad) {
                        // labels 0..4: maxDCG = sum_L (2^L - 1)(P[e_L] - P[e_{L+1}]) = 8 P[e4] + 4 P[e3] + 2 P[e2] + P[e1]
                        const int e4 = packed >> 24;
                        const int e3 = e4 + ((packed >> 16) & 255);
                        const int e2 = e3 + ((packed >> 8) & 255);
                        const int e1 = e2 + (packed & 255);
                        maxdcg = fmaf(8.0f, ctl.disc[e4], fmaf(4.0f, ctl.disc[e3], fmaf(2.0f, ctl.disc[e2], ctl.disc[e1])));
                        gain = (float)((1 << lvl) - 1);
                    } else {
                        // other labels: a doc's ideal rank = how many clamped labels are larger (equal ones: by position)
                        sh.part[0][gt] = valid ? yc : -1.0f;
                        named_bar_sync(bar, 128);
                        const float mine = sh.part[0][gt];
                        int rank = 0;
                        for (int j = 0; j < RF_DOCS; ++j) {
                            const float o = sh.part[0][j];
                            rank += (o > mine || (o == mine && j < gt)) ? 1 : 0;
                        }
                        float term = (mine >= 0.f) ? gain_exp2(mine) / log2f(2.0f + (float)rank) : 0.f;
                        term = warp_sum(term);
                        if (lane == 0) sh.red[w] = term;
                        named_bar_sync(bar, 128);
                        maxdcg = (sh.red[0] + sh.red[1]) + (sh.red[2] + sh.red[3]);
                        gain = gain_exp2(yc);
                        named_bar_sync(bar, 128);
                    }
                    maxdcg = fmaxf(maxdcg, p.eps);
                    gain = valid ? __fdividef(gain, maxdcg) : 0.f;    // G_i (approxNDCG.py:44)
                    // compacted per-doc quantities
                    const float smin = float_from_key(gmin), smax = float_from_key(gmax);
                    const float range = (smax - smin) * scale;
                    const bool fact = (fabsf(range) < 126.0f) && !clamp_pairs;   // false for inf / NaN ranges too
                    const float smid = 0.5f * smin + 0.5f * smax;
                    if (valid) {
                        if (fact) {
                            const float tc = (sv - smid) * scale;
                            sh.u[dst] = ex2_approx(tc);
                            sh.v[dst] = ex2_approx(-tc);
                        } else {
                            sh.u[dst] = sv;
                        }
                    }
                    if (gt >= m) {                                    // zero tail: p = 0, w = 0 for padding
                        sh.u[gt] = 0.f;
                        sh.v[gt] = 0.f;
                        sh.c[gt] = 0.f;
                    }
                    named_bar_sync(bar, 128);
                    float gsum;
                    if (!fact) {
                        const float si = valid ? sh.u[dst] : 0.f;
                        float sum = 0.f;
                        if (valid) sum = rf_direct_rank_sum(sh.u, m, scale, si, p.eps, clamp_pairs);
                        slate_val = rf_middle(sh, bar, lane, w, valid, dst, gain, coef, sum);
                        gsum = 0.f;
                        if (valid) gsum = rf_direct_grad_sum(sh.u, sh.c, m, scale, si, sh.c[dst], p.eps, clamp_pairs);
                    } else if (m > 96) {
                        gsum = rf_sym_slate<4>(sh, bar, gt, k, m, valid, dst, gain, coef, slate_val);
                    } else {                                          // (no R = 2 copy for short slates: code size)
                        gsum = rf_sym_slate<3>(sh, bar, gt, k, m, valid, dst, gain, coef, slate_val);
                    }
                    ds = valid ? p.alpha * gsum : 0.f;                // d loss / d score of this doc
                }
            }
#if !RF_EARLY_F
            // The group's next forward MMA may go (it then runs while dZ is written, ahead of this slate's gradient job in
            // the tensor pipe).
            mbar_arrive(&ctl.zfree[g]);
#endif
            ds_sum += ds;
            if (gt == 0) {
                p.slate_ndcg[slate] = slate_val;
                val_sum += slate_val;
            }
            // ---- dZ of this doc -> the group's A tile of the gradient MMA (K-major: row = hidden unit, column = doc) -----------
            const int tile = k % RF_NDZ, tuse = k / RF_NDZ;       // tiles go round in slate order
            // A parity wait is only sound within one phase of the barrier.  The tile's barrier is shared with the other groups,
            // so first make sure that the gradient job of this group's OWN previous slate is complete (jobs complete in slate
            // order, and G <= 2 RF_NDZ: every use of the tile but the latest is then complete), then wait for the latest.
            if (gt == 0) RF_STAMP(g, 3, k);
            if (it >= 1) mbar_wait_bounded(&ctl.dzmine[g], (unsigned)(it - 1) & 1u);
            if (tuse >= 1) mbar_wait_bounded(&ctl.dzfree[tile], (unsigned)(tuse - 1) & 1u);
            if (gt == 0) RF_STAMP(g, 4, k);
            unsigned char* dzw = dz + tile * RF_DZ_TILE_BYTES + w * RF_DZ_CHUNK_BYTES;        // this warp's chunk of 32 docs
            float dzv[RF_H];
#pragma unroll
            for (int j = 0; j < RF_H; ++j) {
                const float aj = a[j];
                const float dzj = (ds * ctl.w3[j]) * fmaf(-aj, aj, aj);
                *reinterpret_cast<float*>(dzw + j * 128 + dzoff[j & 7]) = dzj;
                dzv[j] = dzj;
                a[j] = aj * ds;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA's reads
            mbar_arrive(&ctl.dzfull[tile]);
            if (gt == 0) RF_STAMP(g, 5, k);
            acc_bc += warp_transpose_sum(dzv, lane);
            acc_w3 += warp_transpose_sum(a, lane);
            if (gt == 0) RF_STAMP(g, 6, k);
        }
        // the group's fp32 sums -> its own shared arrays (read after the CTA barrier below)
        const float ds_w = warp_sum(ds_sum);
        sh.part[w][lane] = acc_bc;
        sh.part[w][32 + lane] = acc_w3;
        if (lane == 0) sh.part[w][64] = ds_w, sh.part[w][65] = val_sum;       // (val_sum: thread 0 of the group only)
    } else if (warp_id == 4 * G) {
        // ================= loader: two rings, polled (the whole warp walks the loop, one lane issues the copies) =============
        if (elect_one()) {
            mbar_expect_tx(&ctl.wfull, (unsigned)nch * RF_W_CHUNK_BYTES);
            for (int ck = 0; ck < nch; ++ck) tma_load_2d(wsm + ck * RF_W_CHUNK_BYTES, &map_w, ck * RF_CHUNK, 0, &ctl.wfull);
        }
        __syncwarp();
        // the forward copy of a chunk is read again (by the gradient job) some 50 MB of traffic later: keep it in L2
        // until then, and let the second read release it
        uint64_t keep, drop;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(drop));
        const int total = nsl * nch, total_d = nsl * (RF_DOCS / RF_QDOCS);
        int fn = 0, fk = 0, fc = 0, fslot = 0, fuse = 0;          // F ring: chunk counter, slate, chunk, slot, use of the slot
        int dn = 0, dk = 0, dc = 0, dslot = 0, duse = 0;
        unsigned idle = 0;
        while (fn < total || dn < total_d) {
            bool did = false;
#if RF_LOADER_D_FIRST
            for (int once = 0; once < 1; ++once)
            if (fn < total && (fuse == 0 || __all_sync(LTR_FULL_MASK, mbar_test(&ctl.empty_f[fslot], (unsigned)(fuse - 1) & 1u)))) {
#else
            // forward chunks first: they are the HBM reads, and a late one stalls a compute group at its next slate
            while (fn < total && (fuse == 0 || __all_sync(LTR_FULL_MASK, mbar_test(&ctl.empty_f[fslot], (unsigned)(fuse - 1) & 1u)))) {
#endif
                if (elect_one()) {
                    mbar_expect_tx(&ctl.full_f[fslot], RF_SLOT_BYTES);
                    tma_load_2d_hint(fring + fslot * RF_SLOT_BYTES, &map_x, fc * RF_CHUNK, ((int)blockIdx.x + fk * (int)gridDim.x) * RF_DOCS,
                                     &ctl.full_f[fslot], keep);
                }
                __syncwarp();
                ++fn;
                if (++fc == nch) fc = 0, ++fk;
                if (++fslot == RF_NSF) fslot = 0, ++fuse;
                did = true;
            }
            // the gradient copy of a slate goes after its forward copy (it is meant to hit L2): 32 docs x all chunks per slot
#if RF_LOADER_D_FIRST
            while (dn < total_d && (fk > dk || fn == total) &&
                   (duse == 0 || __all_sync(LTR_FULL_MASK, mbar_test(&ctl.empty_d[dslot], (unsigned)(duse - 1) & 1u)))) {
#else
            if (dn < total_d && (fk > dk || fn == total) &&
                (duse == 0 || __all_sync(LTR_FULL_MASK, mbar_test(&ctl.empty_d[dslot], (unsigned)(duse - 1) & 1u)))) {
#endif
                if (elect_one()) {
                    mbar_expect_tx(&ctl.full_d[dslot], (unsigned)nch * RF_QBOX_BYTES);
                    const int row0 = ((int)blockIdx.x + dk * (int)gridDim.x) * RF_DOCS + dc * RF_QDOCS;
                    for (int ck = 0; ck < nch; ++ck)
                        tma_load_2d_hint(dring + dslot * RF_QSLOT_BYTES + ck * RF_QBOX_BYTES, &map_xm, ck * RF_CHUNK, row0,
                                         &ctl.full_d[dslot], drop);
                }
                __syncwarp();
                ++dn;
                if (++dc == RF_DOCS / RF_QDOCS) dc = 0, ++dk;
                if (++dslot == RF_NSD) dslot = 0, ++duse;
                did = true;
            }
            if (did) {
                idle = 0;
            } else {
                __nanosleep(64);                                  // nothing to load: leave the issue slots to the compute groups
                if (++idle > (1u << 24)) __trap();                // a ring that never drains traps instead of hanging the GPU
            }
        }
    } else if (warp_id == 4 * G + 1) {
        // ================= F issuer: forward MMAs, in slate order (the whole warp walks the loop, one lane issues) ==========
        constexpr unsigned IDESC1 = umma_idesc_tf32(128, RF_H, 0, 0);
        mbar_wait_bounded(&ctl.wfull, 0);
        const unsigned wa = smem_u32(wsm), fa = smem_u32(fring);
        int slot = 0, use = 0;
        for (int k = 0; k < nsl; ++k) {
            const int g = k % G, it = k / G;
            if (lane == 0) RF_STAMP(4, 0, k);
            if (it >= 1) mbar_wait_bounded(&ctl.zfree[g], (unsigned)(it - 1) & 1u);
            if (lane == 0) RF_STAMP(4, 1, k);
            tc_fence_after();
            for (int ck = 0; ck < nch; ++ck) {
                mbar_wait_bounded(&ctl.full_f[slot], (unsigned)use & 1u);
                if (lane == 0 && ck == nch - 1) RF_STAMP(4, 2, k);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(fa + (unsigned)slot * RF_SLOT_BYTES, 16, 1024);
                    const uint64_t bdesc = umma_desc_sw128(wa + (unsigned)ck * RF_W_CHUNK_BYTES, 16, 1024);
                    // a k step = 32 bytes = 2 units of the descriptor's start address; straight-line code (a loop here costs
                    // ~100 cycles of single-thread bookkeeping per MMA, and the forward job is on every group's critical path)
                    const int steps = min(4, ksteps - 4 * ck);
                    const unsigned zc = tmem + (unsigned)(g * RF_H);
                    tc_mma_tf32(zc, adesc, bdesc, IDESC1, ck != 0);
                    if (steps == 4) {
                        tc_mma_tf32(zc, adesc + 2, bdesc + 2, IDESC1, 1);
                        tc_mma_tf32(zc, adesc + 4, bdesc + 4, IDESC1, 1);
                        tc_mma_tf32(zc, adesc + 6, bdesc + 6, IDESC1, 1);
                    } else {
                        if (steps > 1) tc_mma_tf32(zc, adesc + 2, bdesc + 2, IDESC1, 1);
                        if (steps > 2) tc_mma_tf32(zc, adesc + 4, bdesc + 4, IDESC1, 1);
                    }
                    tc_commit(&ctl.empty_f[slot]);
                    if (ck == nch - 1) tc_commit(&ctl.zfull[g]);
                }
                __syncwarp();
                if (lane == 0 && ck == nch - 1) RF_STAMP(4, 3, k);
                if (++slot == RF_NSF) slot = 0, ++use;
            }
        }
    } else {
        // ================= D issuer: gradient MMAs, accumulating in slate order ==============================================
        const unsigned idesc2 = umma_idesc_tf32(64, nch * RF_CHUNK, 0, 1);       // N = all feature chunks at once
        const unsigned dza = smem_u32(dz), dra = smem_u32(dring);
        int slot = 0, use = 0;
        for (int k = 0; k < nsl; ++k) {
            const int tile = k % RF_NDZ, tuse = k / RF_NDZ;
            if (lane == 0) RF_STAMP(5, 0, k);
            mbar_wait_bounded(&ctl.dzfull[tile], (unsigned)tuse & 1u);
            if (lane == 0) RF_STAMP(5, 1, k);
            tc_fence_after();
            for (int q = 0; q < RF_DOCS / RF_QDOCS; ++q) {
                mbar_wait_bounded(&ctl.full_d[slot], (unsigned)use & 1u);
                if (lane == 0) RF_STAMP(5, 2 + q, k);
                tc_fence_after();
                if (elect_one()) {
                    // A: dZ^T K-major, doc quarter q = one 4 KB chunk of the tile, 8 docs = 32 bytes (2 address units) inside its
                    // 128-byte rows (rows 32..63 of the M = 64 tile: whatever follows, never read back).  Read ONCE per k step:
                    // the MMA spans all feature chunks (operand reads from shared memory are what bounds this kernel).
                    // B: the X quarter MN-major: feature chunks 4 KB apart (leading-dimension offset), 8 docs = two groups of
                    // 4 rows = 1 KB (64 units).
                    const uint64_t adesc = umma_desc_sw128(dza + (unsigned)tile * RF_DZ_TILE_BYTES + (unsigned)q * RF_DZ_CHUNK_BYTES, 16, 1024);
                    const uint64_t bdesc = umma_desc(dra + (unsigned)slot * RF_QSLOT_BYTES, RF_QBOX_BYTES, 512, 1u);
#pragma unroll
                    for (int s = 0; s < RF_QDOCS / 8; ++s)
                        tc_mma_tf32(tmem_g, adesc + (uint64_t)(s * 2), bdesc + (uint64_t)(s * 64), idesc2, (k > 0 || q > 0 || s > 0));
                    tc_commit(&ctl.empty_d[slot]);
                    if (q == RF_DOCS / RF_QDOCS - 1) {
                        tc_commit(&ctl.dzfree[tile]);
                        tc_commit(&ctl.dzmine[k % G]);
                        if (k == nsl - 1) tc_commit(&ctl.alldone);
                    }
                }
                __syncwarp();
                if (lane == 0 && q == RF_DOCS / RF_QDOCS - 1) RF_STAMP(5, 6, k);
                if (++slot == RF_NSD) slot = 0, ++use;
            }
        }
    }
    // ---- the CTA's accumulators -> global partials --------------------------------------------------------------------
    if (warp_id < 2) {
        // M = 64 accumulator: row r lives in lane (r % 16) + 32 (r / 16): lanes 0..15 of warps 0 and 1 hold rows 0..31
        mbar_wait_bounded(&ctl.alldone, 0);
        tc_fence_after();
        float* mine = p.part + (size_t)blockIdx.x * RF_PART;
        for (int c0 = 0; c0 < RF_N2; c0 += 32) {
            float v[32];
            if (c0 < nch * RF_CHUNK) {
                tmem_ld32(tmem_g + ((unsigned)(warp_id * 32) << 16) + (unsigned)c0, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (lane < 16) {
                const int row = warp_id * 16 + lane;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(mine + (size_t)row * RF_N2 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 2 * RF_H + 2) {
        float t = 0.f;
        for (int g = 0; g < G; ++g) {
            const RfGroup& sg = groups[g];
            if (tid < 2 * RF_H + 1) t += (sg.part[0][tid] + sg.part[1][tid]) + (sg.part[2][tid] + sg.part[3][tid]);
            else t += sg.part[0][tid];                            // sum of slate values: thread 0 of the group
        }
        if (tid < 2 * RF_H) p.part_bias[(size_t)blockIdx.x * (2 * RF_H) + tid] = t;
        else p.part_misc[2 * blockIdx.x + (tid == 2 * RF_H ? 1 : 0)] = t;
    }
    if (warp_id == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(RF_TMEM_COLS) : "memory");
}

// partials of the CTAs, added in a fixed order: gWc [32, F], gbc [32], gw3 [32], gb3 [1], loss [1].
// A block owns 32 consecutive entries; its 8 warps add up every 8th CTA's partial (independent coalesced loads) and the
// eight sums meet in shared memory.
constexpr int RF_RED_WARPS = 8;
__global__ void __launch_bounds__(32 * RF_RED_WARPS) ranker_fused_reduce_kernel(const RankerFusedParams p, float* gWc, float* gbc,
                                                                                float* gw3, float* gb3) {
    __shared__ float sums[RF_RED_WARPS][33];
    __shared__ double dsum[RF_RED_WARPS][2];
    pdl_prologue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nblk_w = RF_PART / 32;                             // blocks over the weight gradient, then 2 over dbc / dw3, then 1
    const int blk = (int)blockIdx.x;
    float t = 0.f;
    if (blk < nblk_w + 2) {
        const float* src = (blk < nblk_w) ? p.part + (size_t)blk * 32 + lane : p.part_bias + (size_t)(blk - nblk_w) * 32 + lane;
        const size_t stride = (blk < nblk_w) ? (size_t)RF_PART : (size_t)(2 * RF_H);
#pragma unroll 4
        for (int g = warp; g < p.grid; g += RF_RED_WARPS) t += src[(size_t)g * stride];
        sums[warp][lane] = t;
    } else {
        double v = 0.0, d = 0.0;
        for (int g = (int)threadIdx.x; g < p.grid; g += 32 * RF_RED_WARPS) v += (double)p.part_misc[2 * g], d += (double)p.part_misc[2 * g + 1];
        v = warp_sum(v);
        d = warp_sum(d);
        if (lane == 0) dsum[warp][0] = v, dsum[warp][1] = d;
    }
    __syncthreads();
    if (warp != 0) return;
    if (blk < nblk_w + 2) {
        float tot = 0.f;
#pragma unroll
        for (int q = 0; q < RF_RED_WARPS; ++q) tot += sums[q][lane];
        if (blk < nblk_w) {
            const int idx = blk * 32 + lane, row = idx / RF_N2, col = idx % RF_N2;
            if (col < p.F) gWc[row * p.F + col] = tot;
        } else if (blk == nblk_w) {
            gbc[lane] = tot;
        } else {
            gw3[lane] = tot;
        }
    } else if (lane == 0) {
        double v = 0.0, d = 0.0;
        for (int q = 0; q < RF_RED_WARPS; ++q) v += dsum[q][0], d += dsum[q][1];
        p.loss[0] = p.loss_kind == 1 ? (float)(-v) : (float)(-v / (double)p.B);
        gb3[0] = (float)d;
    }
}

// ---- the linear part of the net, folded and unfolded on the device (whole-step entry) -----------------------------------------
constexpr int RF_H1 = 64;                                        // l1: F -> 64, l2: 64 -> 32 (architeture/tripleLayer.py:8-10)
// Wc = W2 W1 [32, F], bc = W2 b1 + b2 [32]; one thread per entry, sums in index order, all 64 loads of an entry in flight
__global__ void __launch_bounds__(256) ranker_fold_kernel(const float* W1, const float* b1, const float* W2, const float* b2, int F,
                                                          float* Wc, float* bc) {
    pdl_prologue();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx < RF_H * F) {
        const int j = idx / F, f = idx % F;
        float w1[RF_H1];
#pragma unroll
        for (int h = 0; h < RF_H1; ++h) w1[h] = __ldg(W1 + h * F + f);          // coalesced over f
        float t = 0.f;
#pragma unroll
        for (int h = 0; h < RF_H1; ++h) t = fmaf(__ldg(W2 + j * RF_H1 + h), w1[h], t);
        Wc[idx] = t;
    } else if (idx < RF_H * F + RF_H) {
        const int j = idx - RF_H * F;
        float t = b2[j];
#pragma unroll 8
        for (int h = 0; h < RF_H1; ++h) t = fmaf(W2[j * RF_H1 + h], b1[h], t);
        bc[j] = t;
    }
}
// dW1 = W2^T gWc [64, F], db1 = W2^T gbc [64], dW2 = gWc W1^T + gbc b1^T [32, 64], db2 = gbc [32], dW3 = gw3 [32].
// Blocks [0, nb1): one thread per entry of dW1 (32 terms, loads in flight together); blocks [nb1, nb1 + 256): one WARP per
// entry of dW2 (F terms, coalesced over the features, fixed shuffle tree); last block: the three small vectors.
__global__ void __launch_bounds__(256) ranker_unfold_kernel(const float* W1, const float* b1, const float* W2, const float* gWc,
                                                            const float* gbc, const float* gw3, int F, float* gW1, float* gb1,
                                                            float* gW2, float* gb2, float* gW3) {
    pdl_prologue();
    const int nb1 = (RF_H1 * F + 255) / 256;
    const int blk = (int)blockIdx.x, tid = (int)threadIdx.x;
    if (blk < nb1) {
        const int idx = blk * 256 + tid;
        if (idx >= RF_H1 * F) return;
        const int h = idx / F, f = idx % F;
        float g[RF_H];
#pragma unroll
        for (int j = 0; j < RF_H; ++j) g[j] = __ldg(gWc + j * F + f);           // coalesced over f
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < RF_H; ++j) t = fmaf(__ldg(W2 + j * RF_H1 + h), g[j], t);
        gW1[idx] = t;
    } else if (blk < nb1 + RF_H * RF_H1 / 8) {
        const int out = (blk - nb1) * 8 + (tid >> 5), lane = tid & 31;          // 8 warps per block, one entry each
        const int j = out / RF_H1, h = out % RF_H1;
        float t = 0.f;
        for (int f = lane; f < F; f += 32) t = fmaf(__ldg(gWc + j * F + f), __ldg(W1 + h * F + f), t);
        t = warp_sum(t);
        if (lane == 0) gW2[out] = fmaf(gbc[j], b1[h], t);
    } else {
        if (tid < RF_H1) {
            float t = 0.f;
#pragma unroll 8
            for (int j = 0; j < RF_H; ++j) t = fmaf(W2[j * RF_H1 + tid], gbc[j], t);
            gb1[tid] = t;
        }
        if (tid < RF_H) gb2[tid] = gbc[tid], gW3[tid] = gw3[tid];
    }
}

// Measurement probe: what the TMA stream of feature tiles delivers on its own (no MMA, no consumer): every CTA keeps
// `slots` copies in flight over its slates.  mode 0: [128 docs][32 floats] boxes through a tensor map (the F ring's copies),
// mode 1: the same bytes as contiguous 4 KB bulk copies (what a layout-free stream of the same tensor reaches),
// mode 2: [32 docs][all chunks] slots of 4 KB boxes with the 32-byte-atom swizzle (the D ring's copies).
__global__ void __launch_bounds__(32, 1) probe_tma_kernel(const __grid_constant__ CUtensorMap map, const float* X, int B, int F, int slots,
                                                          int mode, int reps) {
    extern __shared__ unsigned char rf_smem_raw[];
    // (aligned by an offset on the __shared__ array itself: a round trip through uintptr_t makes every access a generic LD / ST)
    unsigned char* base = rf_smem_raw + ((1024u - (smem_u32(rf_smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nsl = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nch = (F + RF_CHUNK - 1) / RF_CHUNK;
    const int slate_bytes = RF_DOCS * F * 4;
    const int per = mode == 0 ? nch : mode == 2 ? RF_DOCS / RF_QDOCS : (slate_bytes + 4095) / 4096;        // copies per slate
    const int total = nsl * per * reps;                               // reps > 1: the same slates again (L2-resident runs)
    int slot = 0, use = 0, k = 0, c = 0;
    for (int n = 0; n < total; ++n) {
        if (use >= 1) mbar_wait_bounded(&full[slot], (unsigned)(use - 1) & 1u);
        if (elect_one()) {
            const int slate = (int)blockIdx.x + k * (int)gridDim.x;
            if (mode == 0) {
                mbar_expect_tx(&full[slot], RF_SLOT_BYTES);
                tma_load_2d(base + slot * RF_SLOT_BYTES, &map, c * RF_CHUNK, slate * RF_DOCS, &full[slot]);
            } else if (mode == 2) {
                mbar_expect_tx(&full[slot], (unsigned)nch * RF_QBOX_BYTES);
                for (int ck = 0; ck < nch; ++ck)
                    tma_load_2d(base + slot * RF_QSLOT_BYTES + ck * RF_QBOX_BYTES, &map, ck * RF_CHUNK, slate * RF_DOCS + c * RF_QDOCS,
                                &full[slot]);
            } else {
                const int off = c * 4096, len = min(4096, slate_bytes - off);
                mbar_expect_tx(&full[slot], (unsigned)len);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(base + slot * RF_SLOT_BYTES)),
                             "l"(reinterpret_cast<const unsigned char*>(X) + (size_t)slate * slate_bytes + off), "r"(len),
                             "r"(smem_u32(&full[slot]))
                             : "memory");
            }
        }
        __syncwarp();
        if (++c == per) c = 0, k = (k + 1 == nsl) ? 0 : k + 1;
        if (++slot == slots) slot = 0, ++use;
    }
    for (int i = 0; i < slots && i < total; ++i) {                   // drain: the last copy into every slot
        const int uses = (total - i + slots - 1) / slots;
        mbar_wait_bounded(&full[i], (unsigned)(uses - 1) & 1u);
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
    return grid * (ltr::RF_PART + 2 + 2 * ltr::RF_H) * sizeof(float);
}

namespace ltr {

static unsigned long long* g_rf_trace = nullptr;   // RF_TRACE builds: set by ltr_debug_ranker_trace

template <typename... Args>
static int launch_pdl_args(void (*kernel)(Args...), int grid, int threads, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, args...);
}

// argument checks + the main kernel + the partial reduce, shared by the two entries.  `pdl`: the main kernel is launched
// with programmatic stream serialization (its griddepcontrol.wait then covers the fold kernel before it).
static int triple_core(const char* who, const float* X, const float* y_true, const float* Wc, const float* bc, const float* w3,
                       float b3, const float* b3_ptr, int B, int n, int F, float alpha, float eps, float pad_value,
                       const double* disc_prefix, int disc_len, float* scores, float* slate_ndcg, float* loss, float* gWc, float* gbc,
                       float* gw3, float* gb3, void* scratch, size_t scratch_bytes, cudaStream_t st, bool pdl, int loss_kind = 0) {
    (void)who;
    if (n != RF_DOCS) return set_error(LTR_EINVAL, "fused ranker step: slates of exactly 128 docs (one MMA tile)");
    if (F < 4 || F > RF_MAX_F || F % 4)
        return set_error(LTR_EINVAL, "fused ranker step: 4 <= F <= 160, F a multiple of 4 (16-byte rows for TMA)");
    if (disc_len < RF_DOCS + 1) return set_error(LTR_EINVAL, "fused ranker step: discount table shorter than 129");
    if (loss_kind == 0 && !(alpha > 0.f)) return set_error(LTR_EINVAL, "fused ranker step: alpha must be positive");
    if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wc)) % 16)
        return set_error(LTR_EINVAL, "fused ranker step: X and Wc must be 16-byte aligned");
    if (scratch_bytes < ltr_triple_scratch_bytes(B, n, F)) return set_error(LTR_EINVAL, "fused ranker step: scratch too small");
    int grid = device_sm_count();
    if (grid > B) grid = B;
    CUtensorMap map_x, map_xm, map_w;
    if (int rc = make_map(&map_x, X, (unsigned long long)B * RF_DOCS, F, RF_DOCS, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map(&map_xm, X, (unsigned long long)B * RF_DOCS, F, RF_QDOCS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_map(&map_w, Wc, RF_H, F, RF_H, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    RankerFusedParams p;
    p.y_true = y_true; p.bc = bc; p.w3 = w3; p.b3 = b3; p.b3_ptr = b3_ptr; p.B = B; p.F = F;
    p.alpha = alpha; p.eps = eps; p.pad = pad_value; p.disc_prefix = disc_prefix;
    p.scores = scores; p.slate_ndcg = slate_ndcg; p.loss = loss;
    p.part = reinterpret_cast<float*>(scratch);
    p.part_misc = p.part + (size_t)device_sm_count() * RF_PART;
    p.part_bias = p.part_misc + (size_t)device_sm_count() * 2;
    p.grid = grid;
    p.loss_kind = loss_kind;
    p.trace = g_rf_trace;
    constexpr size_t RF_SMEM = 1024 + (size_t)RF_NSF * RF_SLOT_BYTES + (size_t)RF_NSD * RF_QSLOT_BYTES + RF_W_BYTES + (size_t)RF_NDZ * RF_DZ_TILE_BYTES +
                               (size_t)RF_G * sizeof(RfGroup) + sizeof(RfCtl);
    static_assert(RF_SMEM <= 232448, "ranker_fused_kernel: shared memory of one CTA");
    static_assert((size_t)RF_G * sizeof(RfGroup) >= RF_DZ_CHUNK_BYTES, "rows 32..63 of the last dZ chunk alias the group structs");
    LTR_CUDA(cudaFuncSetAttribute(ranker_fused_kernel<RF_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RF_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(RF_THREADS);
    cfg.dynamicSmemBytes = RF_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    int rc = (int)cudaLaunchKernelEx(&cfg, ranker_fused_kernel<RF_G>, map_x, map_xm, map_w, p);
    if (rc) return set_cuda_error(rc, "fused ranker step (main kernel)");
    rc = launch_pdl_args(ranker_fused_reduce_kernel, RF_PART / 32 + 3, 32 * RF_RED_WARPS, st, p, gWc, gbc, gw3, gb3);
    if (rc) return set_cuda_error(rc, "fused ranker step (reduce)");
    rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "fused ranker step") : 0;
}

}  // namespace ltr

LTR_API int ltr_triple_approx_ndcg_fwdbwd(const float* X, const float* y_true, const float* Wc, const float* bc, const float* w3,
                                          float b3, int B, int n, int F, float alpha, float eps, float pad_value,
                                          const double* disc_prefix, int disc_len, float* scores, float* slate_ndcg, float* loss,
                                          float* gWc, float* gbc, float* gw3, float* gb3, void* scratch, size_t scratch_bytes,
                                          void* workspace, void* stream) {
    using namespace ltr;
    if (!X || !y_true || !Wc || !bc || !w3 || !scores || !slate_ndcg || !loss || !gWc || !gbc || !gw3 || !gb3 || !scratch ||
        !workspace || !disc_prefix || B <= 0)
        return set_error(LTR_EINVAL, "ltr_triple_approx_ndcg_fwdbwd: null pointer or empty batch");
    return triple_core("ltr_triple_approx_ndcg_fwdbwd", X, y_true, Wc, bc, w3, b3, nullptr, B, n, F, alpha, eps, pad_value, disc_prefix,
                       disc_len, scores, slate_ndcg, loss, gWc, gbc, gw3, gb3, scratch, scratch_bytes,
                       reinterpret_cast<cudaStream_t>(stream), false);
}

LTR_API size_t ltr_triple_step_scratch_bytes(int B, int n, int F) {
    if (B <= 0 || F <= 0) return 0;
    // per-CTA partials, then Wc [32, F], bc [32], gWc [32, F], gbc [32], gw3 [32] (each region 16-byte aligned)
    return ltr_triple_scratch_bytes(B, n, F) + (size_t)(2 * ltr::RF_H * F + 3 * ltr::RF_H) * sizeof(float);
}

namespace ltr {
// fold -> fused kernel -> partial reduce -> unfold, for either loss
static int triple_step(const char* who, int loss_kind, const float* X, const float* y_true, const float* W1, const float* b1,
                       const float* W2, const float* b2, const float* W3, const float* b3, int B, int n, int F, float alpha, float eps,
                       float pad_value, const double* disc_prefix, int disc_len, float* scores, float* slate_val, float* loss,
                       float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3, void* scratch, size_t scratch_bytes,
                       void* workspace, void* stream) {
    if (!X || !y_true || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !scores || !slate_val || !loss || !gW1 || !gb1 || !gW2 || !gb2 ||
        !gW3 || !gb3 || !scratch || !workspace || !disc_prefix || B <= 0)
        return set_error(LTR_EINVAL, "fused ranker step: null pointer or empty batch");
    if (F < 4 || F > RF_MAX_F || F % 4) return set_error(LTR_EINVAL, "fused ranker step: 4 <= F <= 160, F a multiple of 4");
    if (scratch_bytes < ltr_triple_step_scratch_bytes(B, n, F)) return set_error(LTR_EINVAL, "fused ranker step: scratch too small");
    if (reinterpret_cast<uintptr_t>(scratch) % 16) return set_error(LTR_EINVAL, "fused ranker step: scratch must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t core = ltr_triple_scratch_bytes(B, n, F);
    float* Wc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + core);
    float* bc = Wc + RF_H * F;
    float* gWc = bc + RF_H;
    float* gbc = gWc + RF_H * F;
    float* gw3 = gbc + RF_H;
    ranker_fold_kernel<<<(RF_H * F + RF_H + 255) / 256, 256, 0, st>>>(W1, b1, W2, b2, F, Wc, bc);
    int rc = (int)cudaGetLastError();
    if (rc) return set_cuda_error(rc, "fused ranker step (fold)");
    rc = triple_core(who, X, y_true, Wc, bc, W3, 0.f, b3, B, n, F, alpha, eps, pad_value, disc_prefix, disc_len, scores, slate_val, loss,
                     gWc, gbc, gw3, gb3, scratch, core, st, true, loss_kind);
    if (rc) return rc;
    rc = launch_pdl_args(ranker_unfold_kernel, (RF_H1 * F + 255) / 256 + RF_H * RF_H1 / 8 + 1, 256, st, W1, b1, W2, (const float*)gWc,
                         (const float*)gbc, (const float*)gw3, F, gW1, gb1, gW2, gb2, gW3);
    if (rc) return set_cuda_error(rc, "fused ranker step (unfold)");
    return 0;
}
}  // namespace ltr

LTR_API int ltr_triple_approx_ndcg_step(const float* X, const float* y_true, const float* W1, const float* b1, const float* W2,
                                        const float* b2, const float* W3, const float* b3, int B, int n, int F, float alpha, float eps,
                                        float pad_value, const double* disc_prefix, int disc_len, float* scores, float* slate_ndcg,
                                        float* loss, float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3,
                                        void* scratch, size_t scratch_bytes, void* workspace, void* stream) {
    return ltr::triple_step("ltr_triple_approx_ndcg_step", 0, X, y_true, W1, b1, W2, b2, W3, b3, B, n, F, alpha, eps, pad_value,
                            disc_prefix, disc_len, scores, slate_ndcg, loss, gW1, gb1, gW2, gb2, gW3, gb3, scratch, scratch_bytes,
                            workspace, stream);
}

LTR_API int ltr_triple_listnet_step(const float* X, const float* y_true, const float* W1, const float* b1, const float* W2,
                                    const float* b2, const float* W3, const float* b3, int B, int n, int F, const double* disc_prefix,
                                    int disc_len, float* scores, float* slate_loss, float* loss, float* gW1, float* gb1, float* gW2,
                                    float* gb2, float* gW3, float* gb3, void* scratch, size_t scratch_bytes, void* workspace,
                                    void* stream) {
    return ltr::triple_step("ltr_triple_listnet_step", 1, X, y_true, W1, b1, W2, b2, W3, b3, B, n, F, 1.0f, 0.f, -1.0f, disc_prefix,
                            disc_len, scores, slate_loss, loss, gW1, gb1, gW2, gb2, gW3, gb3, scratch, scratch_bytes, workspace, stream);
}

// Feature-tile TMA stream on its own (measurement; see probe_tma_kernel): bytes moved = reps * B * 128 * F * 4.
LTR_API int ltr_probe_tma_stream(const float* X, int B, int F, int slots, int mode, int reps, void* stream) {
    using namespace ltr;
    if (!X || B <= 0 || F < 4 || F > RF_MAX_F || F % 4 || slots < 1 || slots > 13 || mode < 0 || mode > 2 || reps < 1 || (mode == 2 && slots > 10))
        return set_error(LTR_EINVAL, "ltr_probe_tma_stream: bad arguments");
    CUtensorMap map;
    if (int rc = mode == 2 ? make_map(&map, X, (unsigned long long)B * RF_DOCS, F, RF_QDOCS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                           : make_map(&map, X, (unsigned long long)B * RF_DOCS, F, RF_DOCS, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    int grid = device_sm_count();
    if (grid > B) grid = B;
    const int smem = 1024 + slots * (mode == 2 ? RF_QSLOT_BYTES : RF_SLOT_BYTES);
    LTR_CUDA(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_tma_kernel<<<grid, 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(map, X, B, F, slots, mode, reps);
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_probe_tma_stream") : 0;
}

#ifdef RF_TRACE
// diagnostics build only (tools/trace_ranker.py): globaltimer stamps of CTA 0 go to `buf` ([6][8][64] u64) from the next call on
LTR_API int ltr_debug_ranker_trace(unsigned long long* buf) {
    ltr::g_rf_trace = buf;
    return 0;
}
#endif
