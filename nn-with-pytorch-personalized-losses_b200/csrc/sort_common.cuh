// K5: CTA-level descending sort of one slate with the tie order of the reference's sorts.
//
// The reference sorts with `torch.sort(descending=True)` / `torch.argsort` on the CPU with the default
// stable=False, which is libstdc++ `std::sort` (introsort: median-of-3 quicksort down to 16 elements, heap
// sort when the depth limit 2*floor(log2 n) is hit, then one insertion sort) over (key, index) pairs with
// the key-only comparator "NaN first, else greater" (SURVEY A.3).  For distinct keys any correct sort
// gives the same permutation; for tied keys the result is whatever that exact sequence of swaps
// leaves behind.  So:
//   1. a bitonic network sorts (key desc, NaN first, index asc) -- the STABLE order -- in shared memory;
//   2. if the slate has tied keys and more than 16 docs (std::sort is an insertion sort, hence
//      stable, up to 16), one thread replays std::sort's exact sequence of moves on a copy in the original order.
// Tie rules: LTR_TIES_STABLE (python sorted(reverse=True), metrics.py:51), LTR_TIES_STDSORT
// (torch CPU), LTR_TIES_REVERSED (np.argsort(kind="stable")[::-1]: tied docs in descending index order).
#pragma once
#include "ltr_common.cuh"

namespace ltr {

constexpr int TIES_STABLE = 0;
constexpr int TIES_STDSORT = 1;
constexpr int TIES_REVERSED = 2;

// key-only comparator of at::native KeyValueCompDesc: a goes before b
__device__ __forceinline__ bool key_before(float a, float b) {
    return (a != a && b == b) || (a > b);
}
__device__ __forceinline__ bool key_equiv(float a, float b) { return !key_before(a, b) && !key_before(b, a); }

// composite order: key (NaN first, then descending), ties by index (ascending, or descending for REVERSED)
template <bool IDX_DESC>
__device__ __forceinline__ bool pair_before(float ka, int ia, float kb, int ib) {
    if (key_before(ka, kb)) return true;
    if (key_before(kb, ka)) return false;
    return IDX_DESC ? (ia > ib) : (ia < ib);
}

// Bitonic sort of len (power of two) (key, index) pairs in shared memory by all threads of the CTA.
template <bool IDX_DESC, bool WARP = false>
__device__ __forceinline__ void block_bitonic_pairs(float* K, int* I, int len) {
#pragma unroll 1
    for (int k = 2; k <= len; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            grp_sync<WARP>();
            for (int t = grp_tid<WARP>(); t < (len >> 1); t += grp_size<WARP>()) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const bool up = ((lo & k) == 0);                 // this run ends up in "before" order
                const float ka = K[lo], kb = K[hi];
                const int ia = I[lo], ib = I[hi];
                const bool b_first = pair_before<IDX_DESC>(kb, ib, ka, ia);
                if (up ? b_first : !b_first) {
                    K[lo] = kb; K[hi] = ka;
                    I[lo] = ib; I[hi] = ia;
                }
            }
        }
    }
    grp_sync<WARP>();
}

// The same network on ONE 64-bit word per entry: (order-preserving key << 32) | tie-break, sorted descending as an
// unsigned integer -- one compare and two 64-bit shared-memory accesses per side instead of the two-array version's
// four loads, NaN-aware float compares and four stores.  key: NaN = 0xffffffff (first), -0 = +0, sentinel entries
// (index >= n) = 0 (last, below -inf); tie-break: ~index (ascending index first) or index (descending).
__device__ __forceinline__ unsigned sort_key_bits(float f) {
    if (f != f) return 0xffffffffu;
    const unsigned b = __float_as_uint(f + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
template <bool WARP = false>
__device__ __forceinline__ void block_bitonic_packed(unsigned long long* P, int len) {
#pragma unroll 1
    for (int k = 2; k <= len; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            grp_sync<WARP>();
            for (int t = grp_tid<WARP>(); t < (len >> 1); t += grp_size<WARP>()) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const unsigned long long a = P[lo], c = P[hi];
                const bool up = ((lo & k) == 0);                 // this run ends up in descending order
                if (up ? (c > a) : (c < a)) {
                    P[lo] = c;
                    P[hi] = a;
                }
            }
        }
    }
    grp_sync<WARP>();
}

// The same network with the entries in REGISTERS: thread t of a CTA of THREADS threads holds the E consecutive entries
// t E .. t E + E - 1 (len = THREADS * E).  A compare-exchange at distance j is local for j < E, a warp shuffle for
// j < 32 E, and goes through the shared scratch P (len words) only for the few stages that cross warps (3 of the 36 stages of
// 256 entries on 128 threads) -- no barrier, no shared-memory round trip and no index arithmetic on the others: ~19 instead of
// ~42 instructions per thread and stage.  The packed words are distinct (index bits), so every correct network gives the same
// result as block_bitonic_packed.
__device__ __forceinline__ unsigned long long bitonic_pick(unsigned long long a, unsigned long long o, bool take_max) {
    return ((o > a) == take_max) ? o : a;
}
// compare-exchanges at distance J inside a thread's E entries; the run's direction may differ per entry (k < E)
template <int J, int E>
__device__ __forceinline__ void bitonic_local(unsigned long long (&v)[E], int i0, int k) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if ((e & J) == 0 && (e | J) < E) {
            const bool up = ((i0 + e) & k) == 0;
            const unsigned long long a = v[e], c = v[e | J];
            const bool sw = up ? (c > a) : (c < a);
            v[e] = sw ? c : a;
            v[e | J] = sw ? a : c;
        }
    }
}
template <int THREADS, int E>
__device__ __forceinline__ void bitonic_stage_regs(unsigned long long* P, unsigned long long (&v)[E], int k, int j, int lane, int i0) {
    if (j < E) {
        if (j == 1) bitonic_local<1, E>(v, i0, k);
        else if (j == 2) bitonic_local<2, E>(v, i0, k);
        else bitonic_local<4, E>(v, i0, k);
    } else if (j < 32 * E) {
        const int m = j / E;                          // partner lane = lane ^ m
        const bool take_max = (((lane & m) == 0) == ((i0 & k) == 0));     // lower entry of a descending run keeps the larger
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = bitonic_pick(v[e], __shfl_xor_sync(LTR_FULL_MASK, v[e], m), take_max);
    } else {
        const bool take_max = (((i0 & j) == 0) == ((i0 & k) == 0));
#pragma unroll
        for (int e = 0; e < E; ++e) P[i0 + e] = v[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = bitonic_pick(v[e], P[(i0 + e) ^ j], take_max);
        __syncthreads();
    }
}
template <int THREADS, int E>
__device__ __forceinline__ void block_bitonic_packed_regs(unsigned long long* P, unsigned long long (&v)[E]) {
    static_assert(E == 1 || E == 2 || E == 4 || E == 8, "entries per thread");
    constexpr int LEN = THREADS * E;
    const int t = threadIdx.x, lane = t & 31, i0 = t * E;
    if constexpr (E <= 2) {
        // straight-line code (28 / 36 stages): no loop counters, no stage-kind branches, shuffle masks as immediates
#pragma unroll
        for (int k = 2; k <= LEN; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) bitonic_stage_regs<THREADS, E>(P, v, k, j, lane, i0);
        }
    } else {
#pragma unroll 1
        for (int k = 2; k <= LEN; k <<= 1) {
#pragma unroll 1
            for (int j = k >> 1; j > 0; j >>= 1) bitonic_stage_regs<THREADS, E>(P, v, k, j, lane, i0);
        }
    }
}

// block_sort_desc's packed path on block_bitonic_packed_regs (K and I adjacent, len = THREADS * E)
template <int THREADS, int E>
__device__ __forceinline__ void block_sort_packed_regs(float* K, int* I, float* K0, int n, bool rev) {
    const float ninf = __int_as_float(0xff800000);
    unsigned long long* P = reinterpret_cast<unsigned long long*>(K);
    const int i0 = threadIdx.x * E;
    unsigned long long v[E];
    float mine[E];
#pragma unroll
    for (int e = 0; e < E; ++e) mine[e] = (i0 + e < n) ? K[i0 + e] : 0.f;
    __syncthreads();                                           // (K is about to become the network's scratch)
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = i0 + e;
        if (i < n) K0[i] = mine[e];
        const unsigned kb = (i < n) ? sort_key_bits(mine[e]) : 0u;
        v[e] = ((unsigned long long)kb << 32) | (unsigned)(rev ? i : ~i);
    }
    block_bitonic_packed_regs<THREADS, E>(P, v);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const unsigned low = (unsigned)v[e];
        const int idx = (int)(rev ? low : ~low);
        K[i0 + e] = (idx < n) ? K0[idx] : ninf;               // sentinels: -inf, index >= n, i.e. last
        I[i0 + e] = idx;
    }
    __syncthreads();
}

// ---- libstdc++ std::sort, replayed by ONE thread on (K, I) in their original order ---------------------
struct StdSortEmu {
    float* K;
    int* I;
    __device__ __forceinline__ bool lt(int a, int b) const { return key_before(K[a], K[b]); }
    __device__ __forceinline__ void swp(int a, int b) {
        const float k = K[a]; K[a] = K[b]; K[b] = k;
        const int i = I[a]; I[a] = I[b]; I[b] = i;
    }
    __device__ void unguarded_linear_insert(int last) {
        const float vk = K[last];
        const int vi = I[last];
        int next = last - 1;
        while (key_before(vk, K[next])) {
            K[last] = K[next]; I[last] = I[next];
            last = next;
            --next;
        }
        K[last] = vk; I[last] = vi;
    }
    __device__ void insertion_sort(int first, int last) {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (lt(i, first)) {
                const float vk = K[i];
                const int vi = I[i];
                for (int j = i; j > first; --j) { K[j] = K[j - 1]; I[j] = I[j - 1]; }
                K[first] = vk; I[first] = vi;
            } else {
                unguarded_linear_insert(i);
            }
        }
    }
    __device__ void adjust_heap(int first, int hole, int len, float vk, int vi) {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (lt(first + child, first + child - 1)) --child;
            K[first + hole] = K[first + child]; I[first + hole] = I[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            K[first + hole] = K[first + child - 1]; I[first + hole] = I[first + child - 1];
            hole = child - 1;
        }
        int parent = (hole - 1) / 2;                              // __push_heap
        while (hole > top && key_before(K[first + parent], vk)) {
            K[first + hole] = K[first + parent]; I[first + hole] = I[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        K[first + hole] = vk; I[first + hole] = vi;
    }
    __device__ void heap_sort(int first, int last) {              // __partial_sort(first, last, last)
        const int len = last - first;
        if (len >= 2) {
            for (int parent = (len - 2) / 2;; --parent) {         // __make_heap
                adjust_heap(first, parent, len, K[first + parent], I[first + parent]);
                if (parent == 0) break;
            }
        }
        for (int end = last; end - first > 1;) {                  // __sort_heap
            --end;
            const float vk = K[end];
            const int vi = I[end];
            K[end] = K[first]; I[end] = I[first];
            adjust_heap(first, 0, end - first, vk, vi);
        }
    }
    __device__ int partition_pivot(int first, int last) {         // __unguarded_partition_pivot
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;           // __move_median_to_first(first, a, b, c)
        if (lt(a, b)) {
            if (lt(b, c)) swp(first, b);
            else if (lt(a, c)) swp(first, c);
            else swp(first, a);
        } else if (lt(a, c)) swp(first, a);
        else if (lt(b, c)) swp(first, c);
        else swp(first, b);
        int lo = first + 1, hi = last;                            // __unguarded_partition(first + 1, last, first)
        for (;;) {
            while (lt(lo, first)) ++lo;
            --hi;
            while (lt(first, hi)) --hi;
            if (!(lo < hi)) return lo;
            swp(lo, hi);
            ++lo;
        }
    }
    // __unguarded_partition_pivot by a whole WARP (all 32 lanes call it with the same arguments; K / I in shared memory).
    // The sequential loop pairs the t-th position from the left whose key does not sort before the pivot with the t-th
    // position from the right whose key does not sort after it, swaps them while they have not met, and both scans only
    // ever look at positions no swap has touched yet -- so the pairs follow from the ORIGINAL keys: ballots number the
    // stoppers of both kinds into two position lists, the first pair that has met ends the loop, all pairs before it are
    // swapped at once.  Same result as partition_pivot move for move (tests: torch CPU sort on heavy ties, and a Python
    // model of both loops); O(range / 32) steps per lane instead of one thread walking the range.
    // lists: 2 n 16-bit entries of scratch (n <= 65536).
    __device__ int partition_pivot_warp(int first, int last, unsigned short* Ll, unsigned short* Rl, int lane) {
        if (lane == 0) {
            const int mid = first + (last - first) / 2;
            const int a = first + 1, b = mid, c = last - 1;       // __move_median_to_first(first, a, b, c)
            if (lt(a, b)) {
                if (lt(b, c)) swp(first, b);
                else if (lt(a, c)) swp(first, c);
                else swp(first, a);
            } else if (lt(a, c)) swp(first, a);
            else if (lt(b, c)) swp(first, c);
            else swp(first, b);
        }
        __syncwarp();
        const float pk = K[first];
        const unsigned below = (1u << lane) - 1u;
        int nl = 0, nr = 0;
        for (int base = first + 1; base < last; base += 32) {
            const int pos = base + lane;
            const bool in = pos < last;
            const float x = in ? K[pos] : 0.f;
            const bool ls = in && !key_before(x, pk);             // stops the scan from the left
            const bool rs = in && !key_before(pk, x);             // stops the scan from the right
            const unsigned bl = __ballot_sync(0xffffffffu, ls), br = __ballot_sync(0xffffffffu, rs);
            if (ls) Ll[nl + __popc(bl & below)] = (unsigned short)pos;
            if (rs) Rl[nr + __popc(br & below)] = (unsigned short)pos;      // ascending; the scan meets them from the end
            nl += __popc(bl);
            nr += __popc(br);
        }
        __syncwarp();
        const int m = nl < nr ? nl : nr;
        int tstar = m;                                            // first pair that has met
        for (int t0 = 0; t0 < m; t0 += 32) {
            const int t = t0 + lane;
            const bool met = t < m && !((int)Ll[t] < (int)Rl[nr - 1 - t]);
            const unsigned bm = __ballot_sync(0xffffffffu, met);
            if (bm) {
                tstar = t0 + __ffs(bm) - 1;
                break;
            }
        }
        for (int t = lane; t < tstar; t += 32) swp((int)Ll[t], (int)Rl[nr - 1 - t]);
        const int lt_pos = tstar < nl ? (int)Ll[tstar] : 0x7fffffff;
        const int r_prev = tstar > 0 ? (int)Rl[nr - tstar] : last;
        __syncwarp();
        return lt_pos < r_prev ? lt_pos : r_prev;
    }
    // sort() by a whole warp: the partitions cooperatively, the small sequential pieces (heap sort at the depth limit,
    // the final insertion sort) by lane 0.  Every lane keeps the same control state (stk may be shared or per lane).
    __device__ int sort_warp(int n, int limit, int* stk, unsigned short* lists, int lane) {
        if (n <= 1) return n;
        int depth0 = 0;
        for (int t = n; t > 1; t >>= 1) ++depth0;                 // __lg(n)
        depth0 *= 2;
        int bound = n;
        int* stk_first = stk;
        int* stk_last = stk + 48;
        int* stk_depth = stk + 96;
        int sp = 0;
        stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = depth0; sp = 1;
        __syncwarp();
        while (sp > 0) {
            --sp;
            const int first = stk_first[sp];
            int last = stk_last[sp], depth = stk_depth[sp];
            __syncwarp();                                         // (shared stack: everyone has read the entry)
            while (last - first > 16) {
                if (depth == 0) {
                    if (lane == 0) heap_sort(first, last);
                    __syncwarp();
                    break;
                }
                --depth;
                const int cut = partition_pivot_warp(first, last, lists, lists + n, lane);
                if (cut < limit) {
                    stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp;
                } else if (cut < bound) {
                    bound = cut;
                }
                last = cut;
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (bound > 16) {                                     // __final_insertion_sort
                insertion_sort(0, 16);
                for (int i = 16; i < bound; ++i) unguarded_linear_insert(i);
            } else {
                insertion_sort(0, bound);
            }
        }
        __syncwarp();
        return bound;
    }
    // Replays std::sort on n elements; only the first `limit` positions of the result are needed.
    // A partition's right part [cut, last) with cut >= limit is never looked at again: introsort's sub-ranges are
    // disjoint, and in the final insertion sort an element can only move left past keys that sort strictly after
    // it, so nothing from such a part can enter a position below cut.  Positions >= the returned bound are left
    // unsorted (bound >= min(limit, n); bound == n when limit >= n: the complete sort).
    // stk: 3 x 48 ints of scratch (local memory for the CTA-level callers, shared memory for the warp-level one)
    __device__ int sort(int n, int limit, int* stk) {
        if (n <= 1) return n;
        int depth0 = 0;
        for (int t = n; t > 1; t >>= 1) ++depth0;                 // __lg(n)
        depth0 *= 2;
        int bound = n;
        // __introsort_loop with an explicit stack (sub-ranges are disjoint, so their order is immaterial)
        int* stk_first = stk;
        int* stk_last = stk + 48;
        int* stk_depth = stk + 96;
        int sp = 0;
        stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = depth0; sp = 1;
        while (sp > 0) {
            --sp;
            const int first = stk_first[sp];
            int last = stk_last[sp], depth = stk_depth[sp];
            while (last - first > 16) {
                if (depth == 0) {
                    heap_sort(first, last);
                    break;
                }
                --depth;
                const int cut = partition_pivot(first, last);
                if (cut < limit) {
                    stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp;
                } else if (cut < bound) {
                    bound = cut;
                }
                last = cut;
            }
        }
        if (bound > 16) {                                         // __final_insertion_sort
            insertion_sort(0, 16);
            for (int i = 16; i < bound; ++i) unguarded_linear_insert(i);
        } else {
            insertion_sort(0, bound);
        }
        return bound;
    }
};

// Sort one slate.  On entry K[0..n) holds the keys in original order.  On exit K/I[0..n) hold the sorted
// keys and the permutation (I[r] = original index of the doc at rank r).  K, I: `len` = next_pow2(n)
// shared-memory entries each; K0 (n floats) is scratch that keeps the original keys for the replay.
// `tie_limit`: the caller only reads the first tie_limit entries of the result (callers whose trailing entries
// are -inf padding that no output depends on pass the number of real docs; NDCG@k passes kmax + 1).  Only ties
// among those entries trigger the replay, and after a replay entries at and beyond tie_limit may be in any order.
template <bool WARP = false, int FAST_THREADS = 0>
__device__ __forceinline__ void block_sort_desc(float* K, int* I, float* K0, int n, int len, int tie_rule,
                                                int tie_limit = 1 << 30, int* warp_stack = nullptr, bool known_tie = false) {
    const float ninf = __int_as_float(0xff800000);
    if (known_tie && tie_rule == TIES_STDSORT && n > 16) {
        // the caller has seen equal keys among the first tie_limit ranks: the network's result would be thrown away
        const int lim = tie_limit < n ? tie_limit : n;
        for (int i = grp_tid<WARP>(); i < n; i += grp_size<WARP>()) I[i] = i;
        grp_sync<WARP>();
        if (grp_tid<WARP>() < 32) {                              // one warp replays std::sort (K0 is free: partition scratch)
            StdSortEmu emu{K, I};
            if constexpr (WARP) {
                emu.sort_warp(n, lim, warp_stack, reinterpret_cast<unsigned short*>(K0), grp_tid<WARP>());
            } else {
                int stk[3 * 48];
                emu.sort_warp(n, lim, stk, reinterpret_cast<unsigned short*>(K0), grp_tid<WARP>());
            }
        }
        grp_sync<WARP>();
        return;
    }
    // FAST_THREADS = the CTA's thread count (compile time): slates of 1, 2, 4 or 8 entries per thread are sorted in registers
    bool fast = false;
    if constexpr (!WARP && FAST_THREADS > 0) {
        fast = reinterpret_cast<void*>(I) == reinterpret_cast<void*>(K + len) && (reinterpret_cast<uintptr_t>(K) & 7) == 0 &&
               (len == FAST_THREADS || len == 2 * FAST_THREADS || len == 4 * FAST_THREADS || len == 8 * FAST_THREADS);
    }
    if (fast) {
        if constexpr (!WARP && FAST_THREADS > 0) {
            const bool rev = tie_rule == TIES_REVERSED;
            if (len == FAST_THREADS) block_sort_packed_regs<FAST_THREADS, 1>(K, I, K0, n, rev);
            else if (len == 2 * FAST_THREADS) block_sort_packed_regs<FAST_THREADS, 2>(K, I, K0, n, rev);
            else if (len == 4 * FAST_THREADS) block_sort_packed_regs<FAST_THREADS, 4>(K, I, K0, n, rev);
            else block_sort_packed_regs<FAST_THREADS, 8>(K, I, K0, n, rev);
        }
    } else if (reinterpret_cast<void*>(I) == reinterpret_cast<void*>(K + len) && (reinterpret_cast<uintptr_t>(K) & 7) == 0 &&
        len <= 8 * grp_size<WARP>()) {
        // K and I are adjacent: sort ONE packed 64-bit word per entry, in place over both arrays.  Every thread first
        // holds its (at most 8) entries in registers, so the in-place repacking cannot overwrite an unread entry.
        unsigned long long* P = reinterpret_cast<unsigned long long*>(K);
        const bool rev = tie_rule == TIES_REVERSED;
        float mine[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = grp_tid<WARP>() + e * grp_size<WARP>();
            mine[e] = (i < n) ? K[i] : 0.f;
        }
        grp_sync<WARP>();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = grp_tid<WARP>() + e * grp_size<WARP>();
            if (i < n) K0[i] = mine[e];
            if (i < len) {
                const unsigned kb = (i < n) ? sort_key_bits(mine[e]) : 0u;
                P[i] = ((unsigned long long)kb << 32) | (unsigned)(rev ? i : ~i);
            }
        }
        block_bitonic_packed<WARP>(P, len);                                 // (starts and ends with a barrier)
        int idx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = grp_tid<WARP>() + e * grp_size<WARP>();
            const unsigned low = (i < len) ? (unsigned)P[i] : 0u;
            idx[e] = (int)(rev ? low : ~low);
        }
        grp_sync<WARP>();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = grp_tid<WARP>() + e * grp_size<WARP>();
            if (i < len) {
                K[i] = (idx[e] < n) ? K0[idx[e]] : ninf;              // sentinels: -inf, index >= n, i.e. last
                I[i] = idx[e];
            }
        }
        grp_sync<WARP>();
    } else {
    for (int i = grp_tid<WARP>(); i < len; i += grp_size<WARP>()) {
        if (i < n) K0[i] = K[i];
        else K[i] = ninf;
        I[i] = i;                                                 // sentinels: -inf with index >= n, i.e. last
    }
    grp_sync<WARP>();
    if (tie_rule == TIES_REVERSED) {
        // sentinels must stay last among -inf keys although indices now sort descending
        for (int i = n + grp_tid<WARP>(); i < len; i += grp_size<WARP>()) I[i] = -1 - i;
        block_bitonic_pairs<true, WARP>(K, I, len);
    } else {
        block_bitonic_pairs<false, WARP>(K, I, len);
    }
    }
    if (tie_rule != TIES_STDSORT || n <= 16) return;
    int tie = 0;
    const int lim = tie_limit < n ? tie_limit : n;
    for (int i = grp_tid<WARP>(); i + 1 < lim; i += grp_size<WARP>()) tie |= key_equiv(K[i], K[i + 1]) ? 1 : 0;
    if (!grp_or<WARP>(tie)) return;
    grp_sync<WARP>();
    for (int i = grp_tid<WARP>(); i < n; i += grp_size<WARP>()) { K[i] = K0[i]; I[i] = i; }      // back to the original order
    grp_sync<WARP>();
    if (grp_tid<WARP>() < 32) {                                  // one warp replays std::sort; K0 (restored from above) is its scratch
        StdSortEmu emu{K, I};
        // entries at and beyond `lim` may stay unsorted
        if constexpr (WARP) {
            emu.sort_warp(n, lim, warp_stack, reinterpret_cast<unsigned short*>(K0), grp_tid<WARP>());   // (no stack frame in the kernel)
        } else {
            int stk[3 * 48];
            emu.sort_warp(n, lim, stk, reinterpret_cast<unsigned short*>(K0), grp_tid<WARP>());
        }
    }
    grp_sync<WARP>();
}

}  // namespace ltr
