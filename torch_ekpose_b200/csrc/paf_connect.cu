// paf_connect.cu -- stage 4 (PAF line-integral scoring of every candidate limb pair) and the
// first half of stage 5 (per-limb sort + greedy bipartite assignment), one block per
// (limb, image).  Replaces /root/reference/lib/pafprocess/pafprocess.cpp:46-125 and
// get_paf_vectors / roundpaf / comp_candidate (:220-246).
//
// Bit-exactness rules followed here:
//  * every float operation is an explicit round-to-nearest intrinsic (no FMA contraction; the
//    reference is built for baseline x86-64 which has none), IEEE sqrt and division;
//  * roundpaf and criterion2 use the same two double-precision steps as the C++ source;
//  * the ten sample scores are accumulated in sample order by the thread that owns the pair;
//  * candidates are compacted in the reference's push order (a outer, b inner) with an ordered
//    ballot/prefix compaction;
//  * two-pass scoring: a pair needs criterion1 > 6 of its 10 samples above 0.05 (pafprocess.cpp:80,85), so a
//    pair whose four MIDDLE samples (i = 3..6, the most discriminative ones: the ends lie on the true limbs
//    of a and b) all fail can never pass.  Pass 1 evaluates only those four (same float operations as the
//    full evaluation, so the test is exact) and keeps the survivors in pair order; pass 2 runs the full,
//    unmodified evaluation on the survivors only.  In crowded scenes roughly 70 % of the pairs end in pass 1 (from the pass times);
//    the gathers are what the kernel is bound by (fully divergent loads: one L1 tag per lane per load);
//  * sorting (pafprocess.cpp:97): the reference's result depends on HOW std::sort permutes equal
//    scores.  For n <= 16 libstdc++ runs a stable insertion sort, and without ties the order is
//    unique, so all threads rank the candidates in parallel (stable) and look for ties; only when
//    n > 16 AND ties exist does one warp replay libstdc++'s algorithm on the original sequence
//    (introsort: median-of-3 quicksort above 16 elements, heapsort after 2*floor(log2 n) levels,
//    then the final insertion sort), which reproduces the reference's permutation exactly.
#include <stdlib.h>

#include "common.cuh"

namespace ekp {

#ifndef EKP_CONN_THREADS
#define EKP_CONN_THREADS 128
#endif
constexpr int kConnThreads = EKP_CONN_THREADS;
constexpr int kMaxPartLimit = 1024;  // upper bound of the per-context max_part (bitmaps of used peaks are static)

// how stage 4 reads the PAF (gathers from global memory / L2)
enum ConnSrc {
    SRC_GLOBAL = 0,       // one load per channel
    SRC_GLOBAL_VEC2 = 1   // channel-last tensor: both channels of a limb with one 8-byte load
};
// (Staging the limb's two stride-8 planes in shared memory and gathering there was built and measured in round 2 on
// 16 x 1312x736 with 20 / 35 / 50 / 80 people per image: 36.9 vs 26.6, 65.2 vs 50.8, 120 vs 116, 420 vs 442 us against
// these gathers -- the planes leave room for one block per SM, which costs more than the gathers' L1 tag rate until the
// crowd is so heavy that the O(n^2) rank sort dominates either way; removed again, profiles/r2_crowd_planes_vs_gathers.txt.)

struct Sample2 { float x, y; };

// Both channels of a limb at one position of a channel-last tensor.  Every limb's two PAF channels are
// adjacent (ch2 == ch1 + 1, ch1 even: pafprocess.h:16-19), so when the channel count is even and the base
// 8-byte aligned (kVec2) one 8-byte load fetches both -- half the gathers.
template <bool kVec2>
__device__ __forceinline__ Sample2 pair_at(const float* q, int ch1, int ch2) {
    Sample2 r;
    if (kVec2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(q + ch1));
        r.x = v.x; r.y = v.y;
    } else {
        r.x = __ldg(q + ch1);
        r.y = __ldg(q + ch2);
    }
    return r;
}

// `packed` = index of this sample in the pre-gathered list (PAF_PACKED only)
template <int kSrc>
__device__ __forceinline__ Sample2 paf_sample(const PafSource& s, int img, int ly, int lx, int ch1, int ch2, long long packed) {
    constexpr bool kVec2 = kSrc == SRC_GLOBAL_VEC2;
    Sample2 r;
    if (s.mode == PAF_PACKED) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(s.ptr) + packed);
        r.x = v.x; r.y = v.y;
        return r;
    }
    lx = min(max(lx, 0), s.W - 1);  // memory safety only: valid peaks never sample outside
    ly = min(max(ly, 0), s.H - 1);
    if (s.mode == PAF_FULL_HWC) {
        r = pair_at<kVec2>(s.ptr + (((size_t) img * s.H + ly) * s.W + lx) * s.C, ch1, ch2);
    } else if (s.mode == PAF_LO_NEAREST) {
        if (s.layout == EKP_LAYOUT_NHWC) {
            r = pair_at<kVec2>(s.ptr + (((size_t) img * s.h + (ly >> 3)) * s.w + (lx >> 3)) * s.C, ch1, ch2);
        } else {
            r.x = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, ly >> 3, lx >> 3);
            r.y = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, ly >> 3, lx >> 3);
        }
    } else {  // identical arithmetic to materialise_tile (dense_frontend.cu)
        int i0, i1, j0, j1;
        float tx, ty;
        bilin_coord(lx, s.w, i0, i1, tx);
        bilin_coord(ly, s.h, j0, j1, ty);
        Sample2 c00, c01, c10, c11;
        if (s.layout == EKP_LAYOUT_NHWC) {
            const float* b0 = s.ptr + ((size_t) img * s.h + j0) * s.w * s.C;
            const float* b1 = s.ptr + ((size_t) img * s.h + j1) * s.w * s.C;
            c00 = pair_at<kVec2>(b0 + (size_t) i0 * s.C, ch1, ch2); c01 = pair_at<kVec2>(b0 + (size_t) i1 * s.C, ch1, ch2);
            c10 = pair_at<kVec2>(b1 + (size_t) i0 * s.C, ch1, ch2); c11 = pair_at<kVec2>(b1 + (size_t) i1 * s.C, ch1, ch2);
        } else {
            c00.x = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j0, i0); c01.x = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j0, i1);
            c10.x = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j1, i0); c11.x = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j1, i1);
            c00.y = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j0, i0); c01.y = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j0, i1);
            c10.y = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j1, i0); c11.y = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j1, i1);
        }
        r.x = lerp1(lerp1(c00.x, c01.x, tx), lerp1(c10.x, c11.x, tx), ty);
        r.y = lerp1(lerp1(c00.y, c01.y, tx), lerp1(c10.y, c11.y, tx), ty);
    }
    return r;
}

// Pass 1: can the pair still satisfy criterion1 > 6 (pafprocess.cpp:80,85)?  Evaluates samples 3..6 with the
// float operations of score_pair; false when all four are <= 0.05 (then at most 6 of 10 can pass) or the
// two peaks coincide (:66).
template <int kSrc>
__device__ __forceinline__ bool pair_may_pass(const ekp_peak& a, const ekp_peak& b, const PafSource& paf, int img, int ch1, int ch2,
                                              long long packed0) {
    const int dxi = b.x - a.x, dyi = b.y - a.y;
    float vx = (float) dxi, vy = (float) dyi;
    const float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
    if ((double) norm < 1e-12) return false;
    vx = __fdiv_rn(vx, norm);
    vy = __fdiv_rn(vy, norm);
    const float step_x = __fdiv_rn((float) dxi, 10.0f);
    const float step_y = __fdiv_rn((float) dyi, 10.0f);
    Sample2 sv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = 3 + k;
        const int lx = (int) __dadd_rn((double) __fadd_rn((float) a.x, __fmul_rn((float) i, step_x)), 0.5);
        const int ly = (int) __dadd_rn((double) __fadd_rn((float) a.y, __fmul_rn((float) i, step_y)), 0.5);
        sv[k] = paf_sample<kSrc>(paf, img, ly, lx, ch1, ch2, packed0 + i);
    }
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; k++) any |= __fadd_rn(__fmul_rn(vx, sv[k].x), __fmul_rn(vy, sv[k].y)) > 0.05f;
    return any;
}

// pafprocess.cpp:59-94 for one (a, b) pair.  Returns true when the pair becomes a candidate.
template <int kSrc>
__device__ __forceinline__ bool score_pair(const ekp_peak& a, const ekp_peak& b, const PafSource& paf, int img, int ch1, int ch2, int h1,
                                           float& criterion2, long long packed0) {
    const int dxi = b.x - a.x, dyi = b.y - a.y;
    float vx = (float) dxi, vy = (float) dyi;
    const float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
    if ((double) norm < 1e-12) return false;
    vx = __fdiv_rn(vx, norm);
    vy = __fdiv_rn(vy, norm);
    const float step_x = __fdiv_rn((float) dxi, 10.0f);
    const float step_y = __fdiv_rn((float) dyi, 10.0f);
    int lx[10], ly[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {  // roundpaf: (int)((double)float + 0.5)
        lx[i] = (int) __dadd_rn((double) __fadd_rn((float) a.x, __fmul_rn((float) i, step_x)), 0.5);
        ly[i] = (int) __dadd_rn((double) __fadd_rn((float) a.y, __fmul_rn((float) i, step_y)), 0.5);
    }
    Sample2 sv[10];
#pragma unroll
    for (int i = 0; i < 10; i++) sv[i] = paf_sample<kSrc>(paf, img, ly[i], lx[i], ch1, ch2, packed0 + i);  // independent gathers in flight
    float scores = 0.0f;
    int criterion1 = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const float s = __fadd_rn(__fmul_rn(vx, sv[i].x), __fmul_rn(vy, sv[i].y));
        scores = __fadd_rn(scores, s);
        if (s > 0.05f) criterion1++;
    }
    const double penalty = __dsub_rn(__ddiv_rn(__dmul_rn(0.5, (double) h1), (double) norm), 1.0);
    const double mn = penalty < 0.0 ? penalty : 0.0;  // std::min(0.0, penalty)
    criterion2 = (float) __dadd_rn((double) __fdiv_rn(scores, 10.0f), mn);
    return criterion1 > 6 && criterion2 > 0.0f;
}

// ---- libstdc++ std::sort on (score, tag) pairs held in shared memory, comp = score greater ----
struct Cand { float s; unsigned t; };
struct CandArray {
    float* s;
    unsigned* t;
    __device__ __forceinline__ Cand get(int i) const { Cand c; c.s = s[i]; c.t = t[i]; return c; }
    __device__ __forceinline__ void set(int i, Cand c) const { s[i] = c.s; t[i] = c.t; }
    __device__ __forceinline__ void swap(int i, int j) const { Cand a = get(i), b = get(j); set(i, b); set(j, a); }
    __device__ __forceinline__ bool comp(int i, int j) const { return s[i] > s[j]; }
};

__device__ void sort_push_heap(const CandArray& A, int first, int hole, int top, Cand value) {
    int parent = (hole - 1) / 2;
    while (hole > top && A.s[first + parent] > value.s) {
        A.set(first + hole, A.get(first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    A.set(first + hole, value);
}
__device__ void sort_adjust_heap(const CandArray& A, int first, int hole, int len, Cand value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (A.comp(first + child, first + child - 1)) child--;
        A.set(first + hole, A.get(first + child));
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        A.set(first + hole, A.get(first + child - 1));
        hole = child - 1;
    }
    sort_push_heap(A, first, hole, top, value);
}
__device__ void sort_heapsort(const CandArray& A, int first, int last) {  // __partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            sort_adjust_heap(A, first, parent, len, A.get(first + parent));
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        const Cand value = A.get(last);
        A.set(last, A.get(first));
        sort_adjust_heap(A, first, 0, last - first, value);
    }
}
// ---- the same algorithm, one range per WARP ------------------------------------------------------
// Control flow, comparisons and element moves of the quicksort phase are exactly libstdc++'s (same
// order), but every scan ("advance while comp holds") inspects 32 elements per step with a ballot; the
// final insertion sort runs one thread per independent range (see std_sort_desc_block).  All 32 lanes call
// these with identical arguments.
__device__ __forceinline__ int lead_true(unsigned m) { return m == 0xffffffffu ? 32 : __ffs(~m) - 1; }

// __unguarded_partition(lo, hi, pivot)
__device__ int warp_partition(const CandArray& A, int lo, int hi, float pivot, int n) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        for (;;) {  // while (comp(lo, pivot)) ++lo;
            const int idx = lo + lane;
            const int run = lead_true(__ballot_sync(0xffffffffu, idx < n && A.s[idx] > pivot));
            lo += run;
            if (run < 32) break;
        }
        --hi;
        for (;;) {  // while (comp(pivot, hi)) --hi;
            const int idx = hi - lane;
            const int run = lead_true(__ballot_sync(0xffffffffu, idx >= 0 && pivot > A.s[idx]));
            hi -= run;
            if (run < 32) break;
        }
        if (!(lo < hi)) return lo;
        if (lane == 0) A.swap(lo, hi);
        __syncwarp();
        ++lo;
    }
}

// `blocks`: scratch for one packed (first << 16 | last) entry per final range, >= n entries (n <= 65535).
// The same partition, 32 swaps at a time, while the two scan fronts are at least 64 elements apart.  The sequential
// loop alternates "advance lo over elements > pivot", "advance hi over elements < pivot", swap, ++lo: inside a window
// of 32 elements per side the k-th element that stops the left scan is therefore swapped with the k-th element that
// stops the right scan, and no position is examined again after its swap.  So the stoppers of both windows are found
// with two ballots on the original values, the first min(nl, nr) pairs are swapped by one lane each, and the fronts
// move exactly where the sequential scan would stand: past a window whose stoppers are used up, or ON the first unused
// stopper of the other (it waits for a partner from the next window).  The windows shrink with the gap (32, 16, 8
// elements); returns with hi - lo < 16 and the caller finishes with warp_partition.
__device__ void warp_partition_wide(const CandArray& A, int& lo, int& hi, float pivot) {
    const int lane = threadIdx.x & 31;
    while (hi - lo >= 16) {
        // two disjoint windows of W elements at the fronts (the argument above holds for any W with hi - lo >= 2 W)
        const int W = hi - lo >= 64 ? 32 : (hi - lo >= 32 ? 16 : 8);
        const bool in = lane < W;
        const unsigned stopL = __ballot_sync(0xffffffffu, in && !(A.s[lo + lane] > pivot));      // comp(first, pivot) fails
        const unsigned stopR = __ballot_sync(0xffffffffu, in && !(pivot > A.s[hi - W + lane]));  // comp(pivot, last) fails
        const int nl = __popc(stopL), nr = __popc(stopR);
        const int pairs = min(nl, nr);
        if (lane < pairs) {
            const int i = lo + (int) __fns(stopL, 0, lane + 1);          // lane-th stopper from the left
            const int j = hi - W + (int) __fns(stopR, 31, -(lane + 1));  // lane-th stopper from the right
            A.swap(i, j);
        }
        __syncwarp();
        const int lo0 = lo, hi0 = hi;
        lo = nl > pairs ? lo0 + (int) __fns(stopL, 0, pairs + 1) : lo0 + W;
        hi = nr > pairs ? hi0 - W + (int) __fns(stopR, 31, -(pairs + 1)) + 1 : hi0 - W;
    }
}

// ---- ... and by ALL WARPS of the block -----------------------------------------------------------------
// __introsort_loop only ever recurses into disjoint ranges, so the order in which the ranges are finished does not
// change the result: the ranges still to be partitioned sit in an append-only queue, every warp takes the next
// ticket, partitions its range (pushing the right part, keeping the left, like the sequential loop), and ranges of
// <= 16 elements go to the list of the final insertion sort.  The critical path is one root-to-leaf chain of
// partitions (n + n/2 + n/4 ...) instead of all of them one after the other.
struct ReplayWork {
    unsigned* range;   // queue: (first << 16 | last), 0 = not published yet; n <= 65535
    unsigned* depth;   // queue: depth limit left for that range
    int cap;           // queue entries (>= pushes + number of warps)
    unsigned* blocks;  // [n] per element: the range (packed like `range`) the final insertion sort handles it in, 0 = none
    unsigned* tmp;     // [n] scratch of the final permutation (may alias range / depth: the queue is idle by then)
    int* ctl;          // [4]: head (next ticket), tail (next free entry), pending (ranges pushed and not finished)
};

// a range of 2..16 elements the quicksort leaves to the final insertion sort: every element learns its range
__device__ __forceinline__ void replay_add_block(const ReplayWork& W, int first, int last) {
    const int lane = threadIdx.x & 31;
    if (lane < last - first) W.blocks[first + lane] = ((unsigned) first << 16) | (unsigned) last;
}
__device__ __forceinline__ void replay_push(const ReplayWork& W, int first, int last, int depth) {
    __threadfence_block();  // this warp's swaps inside [first, last) before the range is published
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&W.ctl[2], 1);
        const int t = atomicAdd(&W.ctl[1], 1);
        W.depth[t] = (unsigned) depth;
        __threadfence_block();
        reinterpret_cast<volatile unsigned*>(W.range)[t] = ((unsigned) first << 16) | (unsigned) last;
    }
}

// one range of the introsort loop, by one warp (all 32 lanes call this with identical arguments)
__device__ void replay_range(const CandArray& A, int n, int first, int last, int depth, const ReplayWork& W) {
    const int lane = threadIdx.x & 31;
    while (last - first > 16) {
        if (depth == 0) {  // heapsort fallback: never reached by real scenes, kept serial
            if (lane == 0) sort_heapsort(A, first, last);
            __syncwarp();
            return;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        if (lane == 0) {  // __move_median_to_first(first, first+1, mid, last-1)
            const int a = first + 1, b = mid, c = last - 1;
            if (A.comp(a, b)) {
                if (A.comp(b, c)) A.swap(first, b);
                else if (A.comp(a, c)) A.swap(first, c);
                else A.swap(first, a);
            } else if (A.comp(a, c)) A.swap(first, a);
            else if (A.comp(b, c)) A.swap(first, c);
            else A.swap(first, b);
        }
        __syncwarp();
        int plo = first + 1, phi = last;
        const float pivot = A.s[first];
        warp_partition_wide(A, plo, phi, pivot);
        const int cut = warp_partition(A, plo, phi, pivot, n);
        if (last - cut > 16) replay_push(W, cut, last, depth);   // __introsort_loop(cut, last, depth)
        else if (last - cut > 1) replay_add_block(W, cut, last);
        last = cut;
    }
    if (last - first > 1) replay_add_block(W, first, last);  // a range the quicksort leaves to the final insertion sort
}

// Called by every thread of the block (blockDim.x a multiple of 32); A may live in shared or global memory.
__device__ void std_sort_desc_block(const CandArray& A, int n, const ReplayWork& W) {
    if (n <= 0) return;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < W.cap; i += blockDim.x) W.range[i] = 0u;
    for (int i = threadIdx.x; i < n; i += blockDim.x) W.blocks[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        int lg = 0;
        for (int v = n; v > 1; v >>= 1) lg++;
        W.depth[0] = (unsigned) (2 * lg);
        W.range[0] = (unsigned) n;  // (0 << 16) | n
        W.ctl[0] = 0; W.ctl[1] = 1; W.ctl[2] = 1; W.ctl[3] = 0;
    }
    __syncthreads();
    for (;;) {
        unsigned rng = 0u, dep = 0u;
        if (lane == 0) {
            const int t = atomicAdd(&W.ctl[0], 1);
            if (t < W.cap) {
                volatile unsigned* q = W.range;
                for (;;) {
                    rng = q[t];
                    if (rng) break;
                    if (*reinterpret_cast<volatile int*>(&W.ctl[2]) == 0) { rng = q[t]; break; }  // nothing will ever be pushed again
                    __nanosleep(100);  // idle warps must not take shared-memory and issue slots from the working ones
                }
                if (rng) { __threadfence_block(); dep = reinterpret_cast<volatile unsigned*>(W.depth)[t]; }
            }
        }
        rng = __shfl_sync(0xffffffffu, rng, 0);
        dep = __shfl_sync(0xffffffffu, dep, 0);
        if (!rng) break;
        __threadfence_block();
        replay_range(A, n, (int) (rng >> 16), (int) (rng & 0xffffu), (int) dep, W);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicSub(&W.ctl[2], 1);
    }
    __syncthreads();
    // __final_insertion_sort (__insertion_sort on the first 16, then __unguarded_insertion_sort): every element moves left
    // past the elements it is greater than (strict comparison: equal elements keep their order).  After the partitioning
    // above the array is a sequence of ranges of <= 16 elements (or heap-sorted ones) with  left range >= pivot >= right
    // range, so no element ever crosses into the range on its left and the pass is a STABLE sort of every range by
    // itself: one thread per ELEMENT counts the elements of its range that end up before it (greater score, or equal
    // score and earlier position); then all elements move at once -- the same permutation without the dependent
    // load-store chain of a sequential insertion sort.
    // pass 1: every element's destination (its own word of `blocks` is overwritten with it; nobody else reads that word)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int dst = i;
        if (W.blocks[i]) {
            const int first = (int) (W.blocks[i] >> 16), last = (int) (W.blocks[i] & 0xffffu);
            const float v = A.s[i];
            dst = first;
            for (int j = first; j < last; j++) {
                const float sj = A.s[j];
                dst += (sj > v) || (sj == v && j < i);
            }
        }
        W.blocks[i] = (unsigned) dst;
    }
    __syncthreads();
    // pass 2: permute through `tmp` (the work queue's memory, idle by now), scores then tags
    for (int i = threadIdx.x; i < n; i += blockDim.x) W.tmp[W.blocks[i]] = __float_as_uint(A.s[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) A.s[i] = __uint_as_float(W.tmp[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) W.tmp[W.blocks[i]] = A.t[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) A.t[i] = W.tmp[i];
    __syncthreads();
}
__host__ __device__ inline int replay_queue_cap(int n, int nthreads) { return n / 16 + 2 + nthreads / 32; }

#ifdef EKP_CONN_PROFILE  // tools/ only: per-phase time of the slowest block and summed over blocks (ns)
__device__ unsigned long long g_conn_prof[16];
__device__ __forceinline__ unsigned long long prof_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PROF_MARK(k) do { if (threadIdx.x == 0) { const unsigned long long _t = prof_now(); atomicAdd(&g_conn_prof[k], _t - prof_t); atomicMax(&g_conn_prof[8 + k], _t - prof_t); prof_t = _t; } } while (0)
extern "C" int ekp_debug_conn_profile(unsigned long long* out16, int reset) {
    cudaMemcpyFromSymbol(out16, g_conn_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_conn_prof, z, sizeof(z)); }
    return 0;
}
#else
#define PROF_MARK(k) do { } while (0)
#endif

// Ordered compaction step shared by both passes: every thread of the block contributes `flag`; returns the
// number of flagged threads before this one plus `base`, and advances `base` by the block's total (identical
// in every thread).  Two block barriers.
template <int kT>
__device__ __forceinline__ int ordered_slot(bool flag, int* sWarpCnt, int& base) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned mask = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) sWarpCnt[warp] = __popc(mask);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < kT / 32; k++) {
        const int c = sWarpCnt[k];
        if (k < warp) before += c;
        all += c;
    }
    const int pos = base + before + __popc(mask & ((1u << lane) - 1u));
    base += all;
    __syncthreads();
    return pos;
}

// ---- few pairs (every scene but a crowd): ten lanes per pair, one sample each -----------------------------------------
// A thread that owns a whole pair has its 10 x (2..8) gathers serialised by its registers (several round trips to L2);
// with a handful of pairs per block most threads would idle meanwhile.  Here a warp takes three pairs, lane 10 g + i
// evaluates sample i of pair g with the same float operations as score_pair, and the group's first lane adds the ten
// values in sample order (warp shuffles; pafprocess.cpp:78 is a sequential float sum) and applies both criteria: one
// round trip to memory per three pairs per warp.  Candidates are compacted in pair order as everywhere else.
template <int kSrc, int kT>
__device__ __forceinline__ int score_pairs_by_sample(const PafSource& paf, const ekp_peak* __restrict__ sA, const ekp_peak* __restrict__ sB,
                                                     int nA, int nB, int img, int ch1, int ch2, int h1, long long packed_base, int max_cand,
                                                     float* __restrict__ sScore, unsigned* __restrict__ sTag, int* sWarpCnt) {
    const unsigned FULL = 0xffffffffu;
    const int npairs = nA * nB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / 10, i = lane - 10 * g;  // lanes 30, 31: g == 3, idle
    constexpr int kPairsPerIter = 3 * (kT / 32);
    int total = 0;
    for (int base = 0; base < npairs; base += kPairsPerIter) {
        const int pidx = base + 3 * warp + g;
        const bool live = g < 3 && pidx < npairs;
        float s = 0.f, vnorm = 0.f;
        int ia = 0, ib = 0;
        bool degenerate = true;
        if (live) {
            ia = pidx / nB;
            ib = pidx - ia * nB;
            const ekp_peak a = sA[ia], b = sB[ib];
            const int dxi = b.x - a.x, dyi = b.y - a.y;
            float vx = (float) dxi, vy = (float) dyi;
            vnorm = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
            degenerate = (double) vnorm < 1e-12;   // pafprocess.cpp:66
            if (!degenerate) {
                vx = __fdiv_rn(vx, vnorm);
                vy = __fdiv_rn(vy, vnorm);
                const float step_x = __fdiv_rn((float) dxi, 10.0f), step_y = __fdiv_rn((float) dyi, 10.0f);
                const int lx = (int) __dadd_rn((double) __fadd_rn((float) a.x, __fmul_rn((float) i, step_x)), 0.5);  // roundpaf
                const int ly = (int) __dadd_rn((double) __fadd_rn((float) a.y, __fmul_rn((float) i, step_y)), 0.5);
                const Sample2 sv = paf_sample<kSrc>(paf, img, ly, lx, ch1, ch2, (packed_base + pidx) * 10 + i);
                s = __fadd_rn(__fmul_rn(vx, sv.x), __fmul_rn(vy, sv.y));
            }
        }
        const unsigned above = __ballot_sync(FULL, live && !degenerate && s > 0.05f);
        float scores = 0.0f;
#pragma unroll
        for (int j = 0; j < 10; j++) scores = __fadd_rn(scores, __shfl_sync(FULL, s, min(10 * g, 20) + j));  // sample order 0..9
        bool pass = false;
        float crit = 0.f;
        if (live && i == 0 && !degenerate) {
            const int criterion1 = __popc(above & (0x3ffu << (10 * g)));
            const double penalty = __dsub_rn(__ddiv_rn(__dmul_rn(0.5, (double) h1), (double) vnorm), 1.0);
            const double mn = penalty < 0.0 ? penalty : 0.0;  // std::min(0.0, penalty)
            crit = (float) __dadd_rn((double) __fdiv_rn(scores, 10.0f), mn);
            pass = criterion1 > 6 && crit > 0.0f;
        }
        const int pos = ordered_slot<kT>(pass, sWarpCnt, total);  // group leaders are in pair order
        if (pass && pos < max_cand) { sScore[pos] = crit; sTag[pos] = ((unsigned) ia << 16) | (unsigned) ib; }
    }
    return total;
}

// ---- stage 4 for one (limb, image): score all nA x nB pairs; candidates end up in pair order (a outer, b inner) in
// sScore / sTag.  Returns the number of candidates (identical in every thread; may exceed max_cand: overflow).
template <int kSrc, int kT>
__device__ __forceinline__ int score_all_pairs(const PafSource& paf, const ekp_peak* __restrict__ sA, const ekp_peak* __restrict__ sB, int nA, int nB, int img, int ch1, int ch2, int h1,
                                               long long packed_base, int max_cand, float* __restrict__ sScore,
                                               unsigned* __restrict__ sTag, unsigned* __restrict__ sTag2, int* sWarpCnt) {
    const int npairs = nA * nB;
    int total = 0;  // candidates so far, identical in every thread
    // Few pairs (every scene but a crowd): one pass, one round trip to memory.  Otherwise pass 1 thins them out.
    const bool two_pass = npairs > 2 * kT;
    const int surv_window = max_cand;  // pairs per pass-1 window: the survivor list lives in sTag2
    for (int win = 0; win < npairs; win += surv_window) {
        const int win_end = min(win + surv_window, npairs);
        int nsurv = 0;  // pass 1: pairs of this window that can still pass, in pair order
        if (!two_pass) {
            nsurv = win_end - win;
            for (int k = threadIdx.x; k < nsurv; k += kT) sTag2[k] = (unsigned) (win + k);
        }
        for (int base = win; two_pass && base < win_end; base += kT) {
            const int pidx = base + threadIdx.x;
            bool keep = false;
            if (pidx < win_end) {
                const int ia = pidx / nB;
                keep = pair_may_pass<kSrc>(sA[ia], sB[pidx - ia * nB], paf, img, ch1, ch2, (packed_base + pidx) * 10);
            }
            const int pos = ordered_slot<kT>(keep, sWarpCnt, nsurv);
            if (keep) sTag2[pos] = (unsigned) pidx;
        }
        __syncthreads();
        for (int base = 0; base < nsurv; base += kT) {  // pass 2: the full evaluation of the survivors
            const int k = base + threadIdx.x;
            bool pass = false;
            float crit = 0.f;
            int ia = 0, ib = 0;
            if (k < nsurv) {
                const int pidx = (int) sTag2[k];
                ia = pidx / nB;
                ib = pidx - ia * nB;
                pass = score_pair<kSrc>(sA[ia], sB[ib], paf, img, ch1, ch2, h1, crit, (packed_base + pidx) * 10);
            }
            const int pos = ordered_slot<kT>(pass, sWarpCnt, total);
            if (pass && pos < max_cand) { sScore[pos] = crit; sTag[pos] = ((unsigned) ia << 16) | (unsigned) ib; }
        }
        __syncthreads();  // the survivor list is rewritten by the next window
    }

    return total;
}

// Dynamic shared memory of one block: [sA, sB: max_part peaks each] [sScore, sTag, sScore2, sTag2: max_cand words each].  max_part / max_cand are capacities of the context
// (ekp_create_ex), reported through EKP_OVF_PART / EKP_OVF_CANDIDATES when a scene exceeds them.
size_t connect_smem_bytes(int max_part, int max_cand) {
    return 2 * sizeof(ekp_peak) * (size_t) max_part + 4 * sizeof(float) * (size_t) max_cand;
}

// one (limb, image): stage 4, sort, greedy assignment; every thread of the block calls it, every thread returns
template <int kSrc, int kT>
__device__ __forceinline__ void connect_limb(const ConnectParams& P, unsigned char* conn_smem) {
    const PafSource& paf = P.paf;
    const int max_part = P.max_part, max_cand = P.max_cand;
    ekp_peak* sA = reinterpret_cast<ekp_peak*>(conn_smem);
    ekp_peak* sB = sA + max_part;
    float* sScore = reinterpret_cast<float*>(sB + max_part);
    unsigned* sTag = reinterpret_cast<unsigned*>(sScore + max_cand);
    float* sScore2 = reinterpret_cast<float*>(sTag + max_cand);
    unsigned* sTag2 = reinterpret_cast<unsigned*>(sScore2 + max_cand);   // pass-1 survivors while scoring, then the ranked tags
    __shared__ int sTies;
    __shared__ int sWarpCnt[kT / 32];
    __shared__ int sReplayCtl[4];
    __shared__ unsigned sUsedA[kMaxPartLimit / 32], sUsedB[kMaxPartLimit / 32];
    const int limb = blockIdx.x, img = blockIdx.y;
    const int pa = kPairs[limb][0], pb = kPairs[limb][1];
    const int ch1 = kPairsNet[limb][0], ch2 = kPairsNet[limb][1];
    const int* po = P.part_off + (size_t) img * 20;
    const int offA = po[pa], offB = po[pb];
    const int nA = min(po[pa + 1] - offA, max_part), nB = min(po[pb + 1] - offB, max_part);
    int* out_n = P.n_conns + (size_t) img * EKP_NUM_LIMB + limb;
    if (nA == 0 || nB == 0) {  // pafprocess.cpp:52-54
        if (threadIdx.x == 0) *out_n = 0;
        return;
    }
#ifdef EKP_CONN_PROFILE
    unsigned long long prof_t = prof_now();
#endif
    const ekp_peak* L = P.line + (size_t) img * P.max_peaks;
    for (int i = threadIdx.x; i < nA; i += kT) sA[i] = L[offA + i];
    for (int i = threadIdx.x; i < nB; i += kT) sB[i] = L[offB + i];
    if (threadIdx.x < kMaxPartLimit / 32) sUsedA[threadIdx.x] = sUsedB[threadIdx.x] = 0u;
    __syncthreads();

    // ---- stage 4: score all nA x nB pairs; candidates end up in pair order (a outer, b inner) --------
    PROF_MARK(0);  // peaks (and planes) staged
    const int npairs = nA * nB;
    long long packed_base = 0;  // PAF_PACKED (one image): where this limb's samples start in the pre-gathered list
    if (paf.mode == PAF_PACKED) {
        packed_base = paf.pair_base[limb];
        if (npairs != paf.pair_base[limb + 1] - paf.pair_base[limb]) {  // the list was laid out for other counts: refuse
            if (threadIdx.x == 0) { *out_n = 0; atomicOr(P.overflow + img, EKP_OVF_BADPEAK); }
            return;
        }
    }
    // Few pairs (every ordinary scene): ten lanes per pair, one round trip to L2.  Many: one thread per pair, two exact passes.
    const int total = npairs <= P.by_sample_max_pairs
                          ? score_pairs_by_sample<kSrc, kT>(paf, sA, sB, nA, nB, img, ch1, ch2, P.h1, packed_base, max_cand, sScore, sTag, sWarpCnt)
                          : score_all_pairs<kSrc, kT>(paf, sA, sB, nA, nB, img, ch1, ch2, P.h1, packed_base, max_cand, sScore, sTag, sTag2, sWarpCnt);
    PROF_MARK(2);  // scoring (both passes)
    // ---- sort (pafprocess.cpp:97).  std::sort's result is only algorithm-dependent in how it
    // permutes EQUAL scores, and for n <= 16 it is a plain (stable) insertion sort.  So: rank every
    // candidate in parallel (stable order) and detect ties; only when n > 16 AND ties exist does
    // the block replay libstdc++'s introsort on the original sequence.
    const int n = min(total, max_cand);
    if (threadIdx.x == 0) {
        sTies = 0;
        if (total > max_cand) atomicOr(P.overflow + img, EKP_OVF_CANDIDATES);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kT) {
        const float s = sScore[i];
        int rank = 0;
        bool tie = false;
        const int n4 = n & ~3;
        for (int j = 0; j < n4; j += 4) {   // four candidates per shared-memory load (sScore is 16-byte aligned)
            const float4 v = *reinterpret_cast<const float4*>(sScore + j);
            rank += (v.x > s) || (v.x == s && j < i);
            rank += (v.y > s) || (v.y == s && j + 1 < i);
            rank += (v.z > s) || (v.z == s && j + 2 < i);
            rank += (v.w > s) || (v.w == s && j + 3 < i);
            tie |= (v.x == s && j != i) || (v.y == s && j + 1 != i) || (v.z == s && j + 2 != i) || (v.w == s && j + 3 != i);
        }
        for (int j = n4; j < n; j++) {
            const float sj = sScore[j];
            rank += (sj > s) || (sj == s && j < i);
            tie |= (sj == s) && (j != i);
        }
        sScore2[rank] = s;
        sTag2[rank] = sTag[i];
        if (tie) sTies = 1;
    }
    __syncthreads();

    PROF_MARK(3);  // rank sort
    const bool replay = n > 16 && sTies;  // uniform over the block
    if (replay) {                          // all warps replay libstdc++'s std::sort on the original sequence
        CandArray A;
        A.s = sScore; A.t = sTag;
        ReplayWork W;  // the ranked copy is not needed when the replay decides the order: its arrays hold the work lists
        W.cap = replay_queue_cap(n, kT);
        W.range = reinterpret_cast<unsigned*>(sScore2);
        W.depth = W.range + W.cap;
        W.blocks = sTag2;
        W.tmp = reinterpret_cast<unsigned*>(sScore2);
        W.ctl = sReplayCtl;
        std_sort_desc_block(A, n, W);
    }
    PROF_MARK(4);  // std::sort replay
    if (threadIdx.x >= 32) return;   // (the warps meet again at the kernel's barrier)
    // ---- greedy assignment, pafprocess.cpp:98-124: walk the sorted candidates, accept one iff neither of its
    // peaks is used yet on this limb.  Warp 0 takes 32 candidates at a time: the lowest lane whose two peaks
    // are still free is the next accepted connection (same order as the sequential walk); its peaks
    // knock out the other lanes' candidates, and the used sets carry over to the next 32.
    const float* srcS = replay ? sScore : sScore2;
    const unsigned* srcT = replay ? sTag : sTag2;
    unsigned* sAcc = replay ? sTag2 : sTag;   // accepted candidates, in acceptance order (the array the sort left free)
    const int lane = threadIdx.x;
    Conn* out = P.conns + ((size_t) img * EKP_NUM_LIMB + limb) * max_part;
    int nc = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int c = c0 + lane;
        const unsigned tag = c < n ? srcT[c] : 0u;
        const int i1 = tag >> 16, i2 = tag & 0xffff;
        bool alive = c < n && !((sUsedA[i1 >> 5] >> (i1 & 31)) & 1u) && !((sUsedB[i2 >> 5] >> (i2 & 31)) & 1u);
        for (;;) {
            const unsigned m = __ballot_sync(0xffffffffu, alive);
            if (!m) break;
            const int leader = __ffs(m) - 1;
            const int l1 = __shfl_sync(0xffffffffu, i1, leader), l2 = __shfl_sync(0xffffffffu, i2, leader);
            if (lane == leader) {
                sAcc[nc] = (unsigned) c;
                sUsedA[i1 >> 5] |= 1u << (i1 & 31);
                sUsedB[i2 >> 5] |= 1u << (i2 & 31);
            }
            if (i1 == l1 || i2 == l2) alive = false;
            nc++;
        }
        __syncwarp();  // the used sets are read by every lane at the top of the next chunk
    }
    // the accepted connections, written by all lanes at once (pafprocess.cpp:117-123)
    for (int k = lane; k < nc; k += 32) {
        const int c = (int) sAcc[k];
        const unsigned tag = srcT[c];
        const int i1 = tag >> 16, i2 = tag & 0xffff;
        Conn cn;
        cn.cid1 = sA[i1].id; cn.cid2 = sB[i2].id; cn.score = srcS[c];
        const float ps1 = sA[i1].score, ps2 = sB[i2].score;  // == peak_infos_line[cid].score when ids are rows
        cn.s_ext = __fadd_rn(ps2, cn.score);
        cn.s_new = __fadd_rn(__fadd_rn(ps1, ps2), cn.score);
        cn.pad0 = cn.pad1 = cn.pad2 = 0;
        out[k] = cn;
    }
    if (lane == 0) *out_n = nc;
    // the score sums the assembly needs per connection, here where 19 x n warps can fetch the peak scores in
    // parallel (one warp per image would pay the two dependent round trips alone)
    __syncwarp();
    for (int k = lane; !paf.ids_are_rows && k < min(nc, max_part); k += 32) {  // process_paf input: ids index the table
        const float sc = out[k].score;
        const float p1 = L[out[k].cid1].score, p2 = L[out[k].cid2].score;
        out[k].s_ext = __fadd_rn(p2, sc);
        out[k].s_new = __fadd_rn(__fadd_rn(p1, p2), sc);
    }
    PROF_MARK(5);  // greedy
#ifdef EKP_CONN_PROFILE
    if (threadIdx.x == 0) { atomicAdd(&g_conn_prof[6], (unsigned long long) n); atomicMax(&g_conn_prof[14], (unsigned long long) n); atomicAdd(&g_conn_prof[7], (unsigned long long) replay); }
#endif
}

// One launch per batch: grid (19 limbs, n images).  (Running the assembly in the same launch -- the block that finishes
// an image's last limb assembles it -- was built and measured in round 2: no gain once a batch is one CUDA graph launch,
// 98.9 vs 100.7 us on 64 x 368x432 and 211.9 vs 214.0 us on 16 crowded 1312x736, and 3 % slower on 256 x 656x368 because
// the assembly's shared memory and registers cost the connect blocks occupancy; removed again, profiles/README.md.)
template <int kSrc, int kT>
__global__ void __launch_bounds__(kT) paf_connect_kernel(const ConnectParams P) {
    extern __shared__ __align__(16) unsigned char conn_smem[];
    connect_limb<kSrc, kT>(P, conn_smem);
}

// ---- host-pointer process_paf: which elements of the caller's paf_mat does stage 4 read? ---------------
// One block per limb of ONE image: for every pair (a outer, b inner) and sample i the element offset of
// channel ch1 at the sample position (roundpaf of pafprocess.cpp:228-233, clamped like paf_sample).  The host
// gathers the two floats at each offset (ch2 == ch1 + 1) and uploads only those; the connect kernel then
// runs in PAF_PACKED mode on identical values.
__global__ void __launch_bounds__(kConnThreads) pair_sample_offsets_kernel(const ekp_peak* __restrict__ line,
                                                                           const int* __restrict__ part_off,
                                                                           const int* __restrict__ pair_base, int H, int W, int C,
                                                                           int max_part, unsigned* __restrict__ offs) {
    const int limb = blockIdx.x;
    const int pa = kPairs[limb][0], pb = kPairs[limb][1];
    const int ch1 = kPairsNet[limb][0];
    const int offA = part_off[pa], offB = part_off[pb];
    const int nA = min(part_off[pa + 1] - offA, max_part), nB = min(part_off[pb + 1] - offB, max_part);
    const int npairs = nA * nB;
    if (npairs != pair_base[limb + 1] - pair_base[limb]) return;  // the host counted differently: it will not use the list
    unsigned* out = offs + (size_t) pair_base[limb] * 10;
    for (int pidx = threadIdx.x; pidx < npairs; pidx += kConnThreads) {
        const int ia = pidx / nB;
        const ekp_peak a = line[offA + ia], b = line[offB + (pidx - ia * nB)];
        const float step_x = __fdiv_rn((float) (b.x - a.x), 10.0f);
        const float step_y = __fdiv_rn((float) (b.y - a.y), 10.0f);
#pragma unroll
        for (int i = 0; i < 10; i++) {
            int lx = (int) __dadd_rn((double) __fadd_rn((float) a.x, __fmul_rn((float) i, step_x)), 0.5);
            int ly = (int) __dadd_rn((double) __fadd_rn((float) a.y, __fmul_rn((float) i, step_y)), 0.5);
            lx = min(max(lx, 0), W - 1);
            ly = min(max(ly, 0), H - 1);
            out[(size_t) pidx * 10 + i] = (unsigned) (((size_t) ly * W + lx) * C + ch1);
        }
    }
}
cudaError_t launch_pair_sample_offsets(const ekp_peak* line, const int* part_off, const int* pair_base, int H, int W, int C,
                                       int max_part, unsigned* offs, cudaStream_t stream) {
    pair_sample_offsets_kernel<<<EKP_NUM_LIMB, kConnThreads, 0, stream>>>(line, part_off, pair_base, H, W, C, max_part, offs);
    return cudaGetLastError();
}

// Test hook: the device replay of libstdc++'s std::sort on caller-supplied scores (one block, as in
// paf_connect_kernel), so that the tie permutation -- including the heapsort fallback, which real scenes never
// reach -- can be compared with the compiled reference's std::sort.  scratch: 2 n + 2 * replay_queue_cap words.
constexpr int kDebugSortThreads = 256;
__global__ void __launch_bounds__(kDebugSortThreads) debug_std_sort_kernel(float* scores, unsigned* tags, int n, unsigned* scratch) {
    __shared__ int ctl[4];
    CandArray A;
    A.s = scores; A.t = tags;
    ReplayWork W;
    W.cap = replay_queue_cap(n, kDebugSortThreads);
    W.blocks = scratch;
    W.tmp = scratch + n;
    W.range = W.tmp + n;
    W.depth = W.range + W.cap;
    W.ctl = ctl;
    std_sort_desc_block(A, n, W);
}
size_t debug_std_sort_scratch_words(int n) { return 2 * (size_t) n + 2 * (size_t) replay_queue_cap(n, kDebugSortThreads); }
cudaError_t launch_debug_std_sort(float* scores, unsigned* tags, int n, unsigned* scratch, cudaStream_t stream) {
    debug_std_sort_kernel<<<1, kDebugSortThreads, 0, stream>>>(scores, tags, n, scratch);
    return cudaGetLastError();
}

// ---- launch ------------------------------------------------------------------------------------------------
// Gathers: 8-byte ones when the tensor is channel-last, even and aligned.  Threads: blocks are latency-bound chains; a
// batch whose 19 x n blocks all fit on the GPU at once (crowded scenes come in small batches) gets twice the threads per
// block, bigger batches keep more blocks resident instead.  (512 threads for those small batches: 16 x 1312x736 with 20 / 35 /
// 50 / 80 people at capacities 256 / 4096: 24.3 / 47.2 / 102 / 393 us against 26.6 / 51.0 / 116 / 443 us, but nothing on the
// bench's crowded configuration (capacities 128 / 1024: 59.4 vs 59.4 us dense, 55.2 vs 53.2 us reference front-end); not kept.)
constexpr size_t kSmemPerSm = 227 * 1024;

template <int kSrc, int kT>
static cudaError_t launch_one(const ConnectParams& P, int n, size_t smem, cudaStream_t stream) {
    paf_connect_kernel<kSrc, kT><<<dim3(EKP_NUM_LIMB, n), kT, smem, stream>>>(P);
    return cudaGetLastError();
}
template <int kSrc>
static cudaError_t launch_src(const ConnectParams& P, int n, size_t smem, int threads, cudaStream_t stream) {
    if (threads == 256) return launch_one<kSrc, 256>(P, n, smem, stream);
    return launch_one<kSrc, 128>(P, n, smem, stream);
}

// per device, once per context: allow the dynamic shared memory this context's capacities ask for
cudaError_t configure_connect(int max_part, int max_cand) {
    const size_t big = connect_smem_bytes(max_part, max_cand);
    if (big > kSmemPerSm) return cudaErrorInvalidValue;
    cudaError_t e = cudaSuccess;
#define EKP_RAISE(S, T) if (e == cudaSuccess) e = raise_dynamic_smem_limit(paf_connect_kernel<S, T>, big)
    EKP_RAISE(SRC_GLOBAL, 128); EKP_RAISE(SRC_GLOBAL, 256);
    EKP_RAISE(SRC_GLOBAL_VEC2, 128); EKP_RAISE(SRC_GLOBAL_VEC2, 256);
#undef EKP_RAISE
    return e;
}

cudaError_t launch_paf_connect(const ConnectParams& P_in, int n, cudaStream_t stream) {
    ConnectParams P = P_in;
    const PafSource& paf = P.paf;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = connect_smem_bytes(P.max_part, P.max_cand);
    const int threads = EKP_NUM_LIMB * n <= 4 * sms ? 2 * kConnThreads : kConnThreads;
    // per-block regimes (same results in both): up to six rounds of ten-lanes-per-pair scoring, beyond that one thread per
    // pair in two exact passes
    static const int env_by_sample = getenv("EKP_BY_SAMPLE_MAX_PAIRS") ? atoi(getenv("EKP_BY_SAMPLE_MAX_PAIRS")) : -1;
    P.by_sample_max_pairs = env_by_sample >= 0 ? env_by_sample : 6 * 3 * (threads / 32);
    // both PAF channels of a limb with one 8-byte load: channel-last tensor, even channel count, aligned base
    const bool channel_last = paf.mode != PAF_PACKED && (paf.mode == PAF_FULL_HWC || paf.layout == EKP_LAYOUT_NHWC);
    const bool vec2 = channel_last && paf.C % 2 == 0 && reinterpret_cast<uintptr_t>(paf.ptr) % 8 == 0;
    return vec2 ? launch_src<SRC_GLOBAL_VEC2>(P, n, smem, threads, stream) : launch_src<SRC_GLOBAL>(P, n, smem, threads, stream);
}

}  // namespace ekp
