// K3: fused forward+backward ListNet (replaces losses/listnet.py:5-16 of the reference) and
// K7: ordinal loss (replaces losses/ordinal.py:9-53).  Both are one-pass streaming kernels bound by
// HBM bandwidth: 12 n bytes per slate for ListNet (read y_true, y_pred; write the gradient).
//
// ListNet, per slate:  p = softmax(y_true), q = softmax(y_pred)
//   loss_b = - sum_i p_i log q_i,                   d loss / d s_k = q_k - p_k            (sum p = 1)
//   apply_sigmoid (listnet.py:13-15): loss_b = - sum_i sig(c_i), c_i = p_i log q_i,
//                                     d loss / d s_k = -sig'(c_k) p_k + q_k sum_i sig'(c_i) p_i
// log q_i is evaluated as (s_i - max) - log sum exp, i.e. without the reference's softmax underflow
// (the two differ only where the reference returns -inf / NaN).  There is no padding mask (listnet.py
// has none).  One warp owns one slate; the whole row lives in registers for n <= 1024.
#include "ltr_host.h"
#include "stream_common.cuh"

namespace ltr {

struct ListnetParams {
    const float* y_true;
    const float* y_pred;
    int B, n, apply_sigmoid;
    float* slate_loss;   // [B] or null
    float* loss;         // [1]
    float* grad;         // [B, n] or null
    double* partials;
    unsigned* counter;
};

template <int ITEMS>
__global__ void __launch_bounds__(STREAM_THREADS) listnet_kernel(const ListnetParams p) {
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * STREAM_WARPS;
    const float ninf = __int_as_float(0xff800000);
    double acc = 0.0;
    for (int b = gwarp; b < p.B; b += nwarps) {
        RowRegs<ITEMS> pt, q;
        pt.load(p.y_true + (size_t)b * p.n, p.n, lane, ninf);
        q.load(p.y_pred + (size_t)b * p.n, p.n, lane, ninf);
        float lp[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) lp[k] = q.v[k];
        softmax_row(pt);
        const float lse = softmax_row(q);
        float part = 0.f;
        if (!p.apply_sigmoid) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const bool in = (lane + 32 * k) < p.n;
                part += in ? pt.v[k] * (lp[k] - lse) : 0.f;
                q.v[k] -= pt.v[k];                                   // gradient
            }
        } else {
            float dsum = 0.f, d[ITEMS];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const bool in = (lane + 32 * k) < p.n;
                const float c = in ? pt.v[k] * (lp[k] - lse) : 0.f;
                const float sg = 1.0f / (1.0f + __expf(-c));
                part += in ? sg : 0.f;
                d[k] = in ? sg * (1.0f - sg) * pt.v[k] : 0.f;
                dsum += d[k];
            }
            dsum = warp_sum(dsum);
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) q.v[k] = fmaf(q.v[k], dsum, -d[k]);
        }
        part = -warp_sum(part);
        if (p.grad) q.store(p.grad + (size_t)b * p.n, p.n, lane);
        if (lane == 0) {
            if (p.slate_loss) p.slate_loss[b] = part;
            acc += (double)part;
        }
    }
    double vals[STREAM_MAX_OUT] = {acc, 0.0, 0.0, 0.0};
    const double scale[STREAM_MAX_OUT] = {1.0, 1.0, 1.0, 1.0};
    stream_finish(vals, 1, p.partials, p.counter, p.loss, scale, nullptr);
}

// Vectorised variant (n % 4 == 0, 16-byte aligned rows): lane l owns the float4s l, l + 32, ... of the row, so a
// 128-doc slate is ONE 128-bit load per lane per array and one 128-bit store for the gradient.  The two row
// maxima are single REDUX instructions on order-preserving integer keys; the three sums the slate needs
//   sy = sum e^(y - my),  ss = sum e^(s - ms),  dot = sum e^(y - my) s
// are reduced together (independent shuffle chains), and  loss_b = lse(s) - dot / sy  (sum p = 1).
template <int V>
__global__ void __launch_bounds__(STREAM_THREADS) listnet_vec_kernel(const ListnetParams p) {
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * STREAM_WARPS;
    const float ninf = __int_as_float(0xff800000);
    const int n4 = p.n >> 2;
    double acc = 0.0;
    pdl_prologue();
    // software pipeline: the next slate's two rows are in flight while this one is reduced
    float4 yn[V], sn[V];
    auto issue_loads = [&](int b) {
        const float4* yrow = reinterpret_cast<const float4*>(p.y_true + (size_t)b * p.n);
        const float4* srow = reinterpret_cast<const float4*>(p.y_pred + (size_t)b * p.n);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int i = lane + 32 * j;
            yn[j] = (i < n4) ? ld_stream4(reinterpret_cast<const float*>(yrow + i)) : make_float4(ninf, ninf, ninf, ninf);
            sn[j] = (i < n4) ? ld_stream4(reinterpret_cast<const float*>(srow + i)) : make_float4(ninf, ninf, ninf, ninf);
        }
    };
    if (gwarp < p.B) issue_loads(gwarp);
    for (int b = gwarp; b < p.B; b += nwarps) {
        float4 y[V], s[V];
#pragma unroll
        for (int j = 0; j < V; ++j) y[j] = yn[j], s[j] = sn[j];
        if (b + nwarps < p.B) issue_loads(b + nwarps);
        float my = ninf, ms = ninf;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            my = fmaxf(my, fmaxf(fmaxf(y[j].x, y[j].y), fmaxf(y[j].z, y[j].w)));
            ms = fmaxf(ms, fmaxf(fmaxf(s[j].x, s[j].y), fmaxf(s[j].z, s[j].w)));
        }
        my = warp_max_redux(my);
        ms = warp_max_redux(ms);
        float sy = 0.f, ss = 0.f, dot = 0.f;
        const float kmy = -my * kLog2e, kms = -ms * kLog2e;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const bool in = (lane + 32 * j) < n4;
            const float4 sv = s[j];
            // e^(x - m) = 2^(x log2e - m log2e); padding lanes give 2^-inf = 0
            y[j].x = ex2_approx(fmaf(y[j].x, kLog2e, kmy)); y[j].y = ex2_approx(fmaf(y[j].y, kLog2e, kmy));
            y[j].z = ex2_approx(fmaf(y[j].z, kLog2e, kmy)); y[j].w = ex2_approx(fmaf(y[j].w, kLog2e, kmy));
            s[j].x = ex2_approx(fmaf(sv.x, kLog2e, kms)); s[j].y = ex2_approx(fmaf(sv.y, kLog2e, kms));
            s[j].z = ex2_approx(fmaf(sv.z, kLog2e, kms)); s[j].w = ex2_approx(fmaf(sv.w, kLog2e, kms));
            sy += (y[j].x + y[j].y) + (y[j].z + y[j].w);
            ss += (s[j].x + s[j].y) + (s[j].z + s[j].w);
            if (in) dot += (y[j].x * sv.x + y[j].y * sv.y) + (y[j].z * sv.z + y[j].w * sv.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sy += __shfl_xor_sync(LTR_FULL_MASK, sy, o);
            ss += __shfl_xor_sync(LTR_FULL_MASK, ss, o);
            dot += __shfl_xor_sync(LTR_FULL_MASK, dot, o);
        }
        const float isy = __fdividef(1.0f, sy), iss = __fdividef(1.0f, ss);
        const float part = (ms + __logf(ss)) - dot * isy;           // -sum p_i log q_i
        if (p.grad) {
            float4* grow = reinterpret_cast<float4*>(p.grad + (size_t)b * p.n);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const int i = lane + 32 * j;
                if (i < n4) {
                    float4 g;
                    g.x = fmaf(s[j].x, iss, -y[j].x * isy); g.y = fmaf(s[j].y, iss, -y[j].y * isy);
                    g.z = fmaf(s[j].z, iss, -y[j].z * isy); g.w = fmaf(s[j].w, iss, -y[j].w * isy);
                    st_stream4(reinterpret_cast<float*>(grow + i), g);
                }
            }
        }
        if (lane == 0) {
            if (p.slate_loss) p.slate_loss[b] = part;
            acc += (double)part;
        }
    }
    double vals[STREAM_MAX_OUT] = {acc, 0.0, 0.0, 0.0};
    const double scale[STREAM_MAX_OUT] = {1.0, 1.0, 1.0, 1.0};
    stream_finish(vals, 1, p.partials, p.counter, p.loss, scale, nullptr);
}

// ---------------------------------------------------------------------------------------------
// ordinal loss: y_pred [B, n, L] probabilities, y_true [B, n]; target_l = [y >= l + 1], label == -1
// marks a padded doc for with_ordinals (its default, ordinal.py:39 does not forward the caller's
// indicator); an entry is masked when its TARGET equals `pad_indicator` (ordinal.py:41).
//   loss = sum over unmasked entries of BCE(pred, target) / #docs with an unmasked entry
// BCE as torch.nn.BCELoss: log terms clamped at -100, gradient (p - t) / max(p (1 - p), 1e-12).
// Two outputs are reduced: [0] the BCE sum, [1] the number of valid docs; the gradient is written
// unscaled and scaled by 1/#valid by `ordinal_scale_kernel` (the count is a batch-wide quantity).
// ---------------------------------------------------------------------------------------------
struct OrdinalParams {
    const float* y_pred;
    const float* y_true;
    long long docs;      // B * n
    int L;
    float pad_indicator;
    float* grad;         // [B, n, L] or null (unscaled on exit of ordinal_kernel)
    float* sums;         // [4]: BCE sum, valid docs, #targets < 0, #inputs outside [0, 1]
    double* partials;
    unsigned* counter;
};

__global__ void __launch_bounds__(STREAM_THREADS) ordinal_kernel(const OrdinalParams p) {
    double bce = 0.0, cnt = 0.0;
    int bad_t = 0, bad_in = 0;
    const long long stride = (long long)gridDim.x * STREAM_THREADS;
    for (long long d = (long long)blockIdx.x * STREAM_THREADS + threadIdx.x; d < p.docs; d += stride) {
        const float y = p.y_true[d];
        const float* pr = p.y_pred + d * p.L;
        float* g = p.grad ? p.grad + d * p.L : nullptr;
        float doc = 0.f;
        int any = 0;
        for (int l = 0; l < p.L; ++l) {
            // torch's BCELoss validates EVERY element before the mask is applied (ordinal.py:44-47)
            bad_in += !(pr[l] >= 0.0f && pr[l] <= 1.0f);
            float t = (y >= (float)(l + 1)) ? 1.0f : 0.0f;
            if (y == -1.0f) t = -1.0f;                               // with_ordinals' own padding value
            bad_t += (t < 0.0f);                                     // a -1 target reaches BCELoss
            const bool masked = (t == p.pad_indicator);
            const float x = pr[l];
            float gl = 0.f;
            if (!masked) {
                const float l1 = fmaxf(logf(x), -100.0f), l0 = fmaxf(logf(1.0f - x), -100.0f);
                doc += -(t * l1 + (1.0f - t) * l0);
                gl = (x - t) / fmaxf(x * (1.0f - x), 1e-12f);
                any = 1;
            }
            if (g) g[l] = gl;
        }
        bce += (double)doc;
        cnt += (double)any;
    }
    double vals[STREAM_MAX_OUT] = {bce, cnt, (double)bad_t, (double)bad_in};
    const double scale[STREAM_MAX_OUT] = {1.0, 1.0, 1.0, 1.0};
    stream_finish(vals, 4, p.partials, p.counter, p.sums, scale, nullptr);
}

// loss[0] = sums[0] / sums[1]; grad *= 1 / sums[1]; status[0] = 1 (a target < 0) | 2 (an input outside [0, 1])
__global__ void ordinal_scale_kernel(const float* sums, float* loss, int* status, float* grad, long long total) {
    const float inv = 1.0f / sums[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        loss[0] = sums[0] / sums[1];
        if (status) status[0] = (sums[2] > 0.f ? 1 : 0) | (sums[3] > 0.f ? 2 : 0);
    }
    if (!grad) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) grad[i] *= inv;
}

static int stream_grid(long long warps_needed) {
    long long g = (warps_needed + STREAM_WARPS - 1) / STREAM_WARPS;
    const long long cap = (long long)device_sm_count() * 8;      // 8 CTAs of 256 threads per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace ltr

LTR_API int ltr_listnet_fwdbwd(const float* y_true, const float* y_pred, int B, int n, int apply_sigmoid,
                               float* slate_loss, float* loss, float* grad, void* workspace, void* stream) {
    using namespace ltr;
    if (!y_true || !y_pred || !loss || !workspace || B <= 0 || n <= 0)
        return set_error(LTR_EINVAL, "ltr_listnet_fwdbwd: null pointer or non-positive B/n");
    if (n > 1024) return set_error(LTR_ETOOBIG, "ltr_listnet_fwdbwd: slate longer than 1024");
    ListnetParams p{y_true, y_pred, B, n, apply_sigmoid, slate_loss, loss, grad, ws_partials(workspace),
                    ws_counter(workspace, LTR_WS_STREAM_COUNTER)};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int grid = stream_grid(B);
    const bool vec = !apply_sigmoid && (n % 4 == 0) && n <= 512 &&
                     ((reinterpret_cast<uintptr_t>(y_true) | reinterpret_cast<uintptr_t>(y_pred) |
                       reinterpret_cast<uintptr_t>(grad)) % 16 == 0);
    if (vec) {
        // one resident wave: the slate assignment is static, a partial second wave would only add a tail
        static int occ[3] = {0, 0, 0};
        const int which = n <= 128 ? 0 : (n <= 256 ? 1 : 2);
        if (occ[which] == 0) {
            if (which == 0) LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], listnet_vec_kernel<1>, STREAM_THREADS, 0));
            else if (which == 1) LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], listnet_vec_kernel<2>, STREAM_THREADS, 0));
            else LTR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], listnet_vec_kernel<4>, STREAM_THREADS, 0));
            if (occ[which] <= 0) occ[which] = 1;
        }
        const int cap = device_sm_count() * occ[which];
        if (grid > cap) grid = cap;
    }
    if (vec) {
        const int rc = n <= 128 ? launch_pdl(listnet_vec_kernel<1>, grid, STREAM_THREADS, 0, st, p)
                     : n <= 256 ? launch_pdl(listnet_vec_kernel<2>, grid, STREAM_THREADS, 0, st, p)
                                : launch_pdl(listnet_vec_kernel<4>, grid, STREAM_THREADS, 0, st, p);
        return rc ? set_cuda_error(rc, "ltr_listnet_fwdbwd") : 0;
    }
    if (n <= 128) listnet_kernel<4><<<grid, STREAM_THREADS, 0, st>>>(p);
    else if (n <= 256) listnet_kernel<8><<<grid, STREAM_THREADS, 0, st>>>(p);
    else if (n <= 512) listnet_kernel<16><<<grid, STREAM_THREADS, 0, st>>>(p);
    else listnet_kernel<32><<<grid, STREAM_THREADS, 0, st>>>(p);
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_listnet_fwdbwd") : 0;
}

LTR_API int ltr_ordinal_fwdbwd(const float* y_pred, const float* y_true, int B, int n, int levels, float pad_indicator,
                               float* loss, int* status, float* grad, void* workspace, void* stream) {
    using namespace ltr;
    if (!y_pred || !y_true || !loss || !workspace || B <= 0 || n <= 0 || levels <= 0)
        return set_error(LTR_EINVAL, "ltr_ordinal_fwdbwd: null pointer or non-positive size");
    float* sums = ws_floats(workspace, LTR_WS_ORDINAL_SUMS);
    OrdinalParams p{y_pred, y_true, (long long)B * n, levels, pad_indicator, grad, sums, ws_partials(workspace),
                    ws_counter(workspace, LTR_WS_STREAM_COUNTER)};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = stream_grid(((long long)B * n + 31) / 32);
    ordinal_kernel<<<grid, STREAM_THREADS, 0, st>>>(p);
    ordinal_scale_kernel<<<grid, STREAM_THREADS, 0, st>>>(sums, loss, status, grad, (long long)B * n * levels);
    const int rc = (int)cudaGetLastError();
    return rc ? set_cuda_error(rc, "ltr_ordinal_fwdbwd") : 0;
}
