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
#include "common.cuh"

namespace ekp {

#ifndef EKP_CONN_THREADS
#define EKP_CONN_THREADS 128
#endif
constexpr int kConnThreads = EKP_CONN_THREADS;
constexpr int kSurvWindow = 2048;  // pairs per pass-1 window (survivor list capacity)

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
template <bool kVec2>
__device__ __forceinline__ Sample2 paf_sample(const PafSource& s, int img, int ly, int lx, int ch1, int ch2, long long packed) {
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
template <bool kVec2>
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
        sv[k] = paf_sample<kVec2>(paf, img, ly, lx, ch1, ch2, packed0 + i);
    }
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; k++) any |= __fadd_rn(__fmul_rn(vx, sv[k].x), __fmul_rn(vy, sv[k].y)) > 0.05f;
    return any;
}

// pafprocess.cpp:59-94 for one (a, b) pair.  Returns true when the pair becomes a candidate.
template <bool kVec2>
__device__ __forceinline__ bool score_pair(const ekp_peak& a, const ekp_peak& b, const PafSource& paf, int img, int ch1,
                                           int ch2, int h1, float& criterion2, long long packed0) {
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
    for (int i = 0; i < 10; i++) sv[i] = paf_sample<kVec2>(paf, img, ly[i], lx[i], ch1, ch2, packed0 + i);  // independent gathers in flight
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
// ---- the same algorithm executed by ONE WARP ---------------------------------------------------
// Control flow, comparisons and element moves of the quicksort phase are exactly libstdc++'s (same
// order), but every scan ("advance while comp holds") inspects 32 elements per step with a ballot; the
// final insertion sort runs one lane per independent range (see std_sort_desc).  All 32 lanes call
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
// stopper of the other (it waits for a partner from the next window).  Returns with hi - lo < 64; the caller finishes
// with warp_partition.
__device__ void warp_partition_wide(const CandArray& A, int& lo, int& hi, float pivot) {
    const int lane = threadIdx.x & 31;
    while (hi - lo >= 64) {
        const unsigned stopL = __ballot_sync(0xffffffffu, !(A.s[lo + lane] > pivot));       // comp(first, pivot) fails
        const unsigned stopR = __ballot_sync(0xffffffffu, !(pivot > A.s[hi - 32 + lane]));  // comp(pivot, last) fails
        const int nl = __popc(stopL), nr = __popc(stopR);
        const int pairs = min(nl, nr);
        if (lane < pairs) {
            const int i = lo + (int) __fns(stopL, 0, lane + 1);           // lane-th stopper from the left
            const int j = hi - 32 + (int) __fns(stopR, 31, -(lane + 1));  // lane-th stopper from the right
            A.swap(i, j);
        }
        __syncwarp();
        const int lo0 = lo, hi0 = hi;
        lo = nl > pairs ? lo0 + (int) __fns(stopL, 0, pairs + 1) : lo0 + 32;
        hi = nr > pairs ? hi0 - 32 + (int) __fns(stopR, 31, -(pairs + 1)) + 1 : hi0 - 32;
    }
}

__device__ void std_sort_desc(const CandArray& A, int n, unsigned* blocks) {
    if (n <= 0) return;
    const int lane = threadIdx.x & 31;
    // __introsort_loop with an explicit stack: the recursion only ever touches disjoint ranges,
    // so the order in which they are finished does not change the result.
    int stk_first[48], stk_last[48], stk_depth[48];
    int sp = 0, nblk = 0;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) lg++;
    stk_first[sp] = 0; stk_last[sp] = n; stk_depth[sp] = 2 * lg; sp++;
    while (sp) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        bool heapsorted = false;
        while (last - first > 16) {
            if (depth == 0) {  // heapsort fallback: never reached by real scenes, kept serial
                if (lane == 0) sort_heapsort(A, first, last);
                __syncwarp();
                heapsorted = true;
                break;
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
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
        if (!heapsorted && last - first > 1) {  // a range the quicksort leaves to the final insertion sort
            if (lane == 0) blocks[nblk] = ((unsigned) first << 16) | (unsigned) last;
            nblk++;
        }
    }
    __syncwarp();
    // __final_insertion_sort (__insertion_sort on the first 16, then __unguarded_insertion_sort): every
    // element moves left past the elements it is greater than.  After the partitioning above the array is
    // a sequence of ranges of <= 16 elements (or heap-sorted ones) with  left range >= pivot >= right range,
    // so no element ever crosses into the range on its left (`val > y` is false for every y there) and the
    // ranges can be insertion-sorted independently: one lane per range, same comparisons and moves per
    // element as the sequential pass, hence the same permutation.
    for (int b = lane; b < nblk; b += 32) {
        const int first = (int) (blocks[b] >> 16), last = (int) (blocks[b] & 0xffffu);
        for (int i = first + 1; i < last; ++i) {
            const Cand val = A.get(i);
            int j = i;
            while (j > first && val.s > A.s[j - 1]) { A.set(j, A.get(j - 1)); --j; }
            if (j != i) A.set(j, val);
        }
    }
    __syncwarp();
}

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

template <bool kVec2, int kT>
__global__ void __launch_bounds__(kT) paf_connect_kernel(const ekp_peak* __restrict__ line,
                                                                   const int* __restrict__ part_off, int max_peaks,
                                                                   const PafSource paf, int h1, Conn* __restrict__ conns,
                                                                   int* __restrict__ n_conns,
                                                                   unsigned* __restrict__ overflow) {
    __shared__ ekp_peak sA[EKP_MAX_PART], sB[EKP_MAX_PART];
    __shared__ float sScore[EKP_MAX_CAND];
    __shared__ unsigned sTag[EKP_MAX_CAND];
    __shared__ float sScore2[EKP_MAX_CAND];
    __shared__ unsigned sTag2[EKP_MAX_CAND];   // pass-1 survivors while scoring, then the ranked tags
    __shared__ int sTies;
    __shared__ int sWarpCnt[kT / 32];
    __shared__ unsigned sUsedA[EKP_MAX_PART / 32], sUsedB[EKP_MAX_PART / 32];
    static_assert(kSurvWindow <= EKP_MAX_CAND, "the survivor list lives in sTag2");
    const int limb = blockIdx.x, img = blockIdx.y;
    const int pa = kPairs[limb][0], pb = kPairs[limb][1];
    const int ch1 = kPairsNet[limb][0], ch2 = kPairsNet[limb][1];
    const int* po = part_off + (size_t) img * 20;
    const int offA = po[pa], offB = po[pb];
    const int nA = min(po[pa + 1] - offA, EKP_MAX_PART), nB = min(po[pb + 1] - offB, EKP_MAX_PART);
    int* out_n = n_conns + (size_t) img * EKP_NUM_LIMB + limb;
    if (nA == 0 || nB == 0) {  // pafprocess.cpp:52-54
        if (threadIdx.x == 0) *out_n = 0;
        return;
    }
#ifdef EKP_CONN_PROFILE
    unsigned long long prof_t = prof_now();
#endif
    const ekp_peak* L = line + (size_t) img * max_peaks;
    for (int i = threadIdx.x; i < nA; i += kT) sA[i] = L[offA + i];
    for (int i = threadIdx.x; i < nB; i += kT) sB[i] = L[offB + i];
    if (threadIdx.x < EKP_MAX_PART / 32) sUsedA[threadIdx.x] = sUsedB[threadIdx.x] = 0u;
    __syncthreads();

    // ---- stage 4: score all nA x nB pairs; candidates end up in pair order (a outer, b inner) --------
    PROF_MARK(0);  // peaks staged
    const int npairs = nA * nB;
    long long packed_base = 0;  // PAF_PACKED (one image): where this limb's samples start in the pre-gathered list
    if (paf.mode == PAF_PACKED) {
        packed_base = paf.pair_base[limb];
        if (npairs != paf.pair_base[limb + 1] - paf.pair_base[limb]) {  // the list was laid out for other counts: refuse
            if (threadIdx.x == 0) { *out_n = 0; atomicOr(overflow + img, EKP_OVF_BADPEAK); }
            return;
        }
    }
    int total = 0;  // candidates so far, identical in every thread
    // Few pairs (every scene but a crowd): one pass, one round trip to memory.  Otherwise pass 1 thins them out.
    const bool two_pass = npairs > 2 * kT;
    for (int win = 0; win < npairs; win += kSurvWindow) {
        const int win_end = min(win + kSurvWindow, npairs);
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
                keep = pair_may_pass<kVec2>(sA[ia], sB[pidx - ia * nB], paf, img, ch1, ch2, (packed_base + pidx) * 10);
            }
            const int pos = ordered_slot<kT>(keep, sWarpCnt, nsurv);
            if (keep) sTag2[pos] = (unsigned) pidx;
        }
        __syncthreads();
        PROF_MARK(1);  // pass 1
        for (int base = 0; base < nsurv; base += kT) {  // pass 2: the full evaluation of the survivors
            const int k = base + threadIdx.x;
            bool pass = false;
            float crit = 0.f;
            int ia = 0, ib = 0;
            if (k < nsurv) {
                const int pidx = (int) sTag2[k];
                ia = pidx / nB;
                ib = pidx - ia * nB;
                pass = score_pair<kVec2>(sA[ia], sB[ib], paf, img, ch1, ch2, h1, crit, (packed_base + pidx) * 10);
            }
            const int pos = ordered_slot<kT>(pass, sWarpCnt, total);
            if (pass && pos < EKP_MAX_CAND) { sScore[pos] = crit; sTag[pos] = ((unsigned) ia << 16) | (unsigned) ib; }
        }
        __syncthreads();  // the survivor list is rewritten by the next window
        PROF_MARK(2);  // pass 2
    }

    // ---- sort (pafprocess.cpp:97).  std::sort's result is only algorithm-dependent in how it
    // permutes EQUAL scores, and for n <= 16 it is a plain (stable) insertion sort.  So: rank every
    // candidate in parallel (stable order) and detect ties; only when n > 16 AND ties exist does
    // one warp replay libstdc++'s introsort on the original sequence.
    const int n = min(total, EKP_MAX_CAND);
    if (threadIdx.x == 0) {
        sTies = 0;
        if (total > EKP_MAX_CAND) atomicOr(overflow + img, EKP_OVF_CANDIDATES);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kT) {
        const float s = sScore[i];
        int rank = 0;
        bool tie = false;
        for (int j = 0; j < n; j++) {
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
    if (threadIdx.x >= 32) return;
    const bool replay = n > 16 && sTies;  // uniform
    if (replay) {                           // warp 0 replays libstdc++'s std::sort on the original sequence
        CandArray A;
        A.s = sScore; A.t = sTag;
        std_sort_desc(A, n, sTag2);  // the ranked copy is not needed when the replay decides the order
    }
    __syncwarp();
    PROF_MARK(4);  // std::sort replay
    // ---- greedy assignment, pafprocess.cpp:98-124: walk the sorted candidates, accept one iff neither of its
    // peaks is used yet on this limb.  Warp 0 takes 32 candidates at a time: the lowest lane whose two peaks
    // are still free is the next accepted connection (same order as the sequential walk); its peaks
    // knock out the other lanes' candidates, and the used sets carry over to the next 32.
    const float* srcS = replay ? sScore : sScore2;
    const unsigned* srcT = replay ? sTag : sTag2;
    const int lane = threadIdx.x;
    Conn* out = conns + ((size_t) img * EKP_NUM_LIMB + limb) * EKP_MAX_PART;
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
                Conn cn;
                cn.cid1 = sA[i1].id; cn.cid2 = sB[i2].id; cn.score = srcS[c];
                const float ps1 = sA[i1].score, ps2 = sB[i2].score;  // == peak_infos_line[cid].score when ids are rows
                cn.s_ext = __fadd_rn(ps2, cn.score);
                cn.s_new = __fadd_rn(__fadd_rn(ps1, ps2), cn.score);
                cn.pad0 = cn.pad1 = cn.pad2 = 0;
                out[nc] = cn;
                sUsedA[i1 >> 5] |= 1u << (i1 & 31);
                sUsedB[i2 >> 5] |= 1u << (i2 & 31);
            }
            if (i1 == l1 || i2 == l2) alive = false;
            nc++;
        }
        __syncwarp();  // the used sets are read by every lane at the top of the next chunk
    }
    if (lane == 0) *out_n = nc;
    // the score sums the assembly needs per connection, here where 19 x n warps can fetch the peak scores in
    // parallel (one warp per image would pay the two dependent round trips alone)
    __syncwarp();
    for (int k = lane; !paf.ids_are_rows && k < min(nc, EKP_MAX_PART); k += 32) {  // process_paf input: ids index the table
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

// ---- host-pointer process_paf: which elements of the caller's paf_mat does stage 4 read? ---------------
// One block per limb of ONE image: for every pair (a outer, b inner) and sample i the element offset of
// channel ch1 at the sample position (roundpaf of pafprocess.cpp:228-233, clamped like paf_sample).  The host
// gathers the two floats at each offset (ch2 == ch1 + 1) and uploads only those; the connect kernel then
// runs in PAF_PACKED mode on identical values.
__global__ void __launch_bounds__(kConnThreads) pair_sample_offsets_kernel(const ekp_peak* __restrict__ line,
                                                                           const int* __restrict__ part_off,
                                                                           const int* __restrict__ pair_base, int H, int W, int C,
                                                                           unsigned* __restrict__ offs) {
    const int limb = blockIdx.x;
    const int pa = kPairs[limb][0], pb = kPairs[limb][1];
    const int ch1 = kPairsNet[limb][0];
    const int offA = part_off[pa], offB = part_off[pb];
    const int nA = min(part_off[pa + 1] - offA, EKP_MAX_PART), nB = min(part_off[pb + 1] - offB, EKP_MAX_PART);
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
                                       unsigned* offs, cudaStream_t stream) {
    pair_sample_offsets_kernel<<<EKP_NUM_LIMB, kConnThreads, 0, stream>>>(line, part_off, pair_base, H, W, C, offs);
    return cudaGetLastError();
}

// Test hook: the device replay of libstdc++'s std::sort on caller-supplied scores (one warp), so that the tie permutation -- including the heapsort fallback, which real scenes never
// reach -- can be compared with the compiled reference's std::sort.
__global__ void debug_std_sort_kernel(float* scores, unsigned* tags, int n, unsigned* scratch) {
    CandArray A;  // one warp, as in paf_connect_kernel
    A.s = scores; A.t = tags;
    std_sort_desc(A, n, scratch);
}
cudaError_t launch_debug_std_sort(float* scores, unsigned* tags, int n, unsigned* scratch, cudaStream_t stream) {
    debug_std_sort_kernel<<<1, 32, 0, stream>>>(scores, tags, n, scratch);
    return cudaGetLastError();
}

cudaError_t launch_paf_connect(const ekp_peak* line, const int* part_off, int max_peaks, const PafSource& paf, int h1,
                               int n, Conn* conns, int* n_conns, unsigned* overflow, cudaStream_t stream) {
    dim3 grid(EKP_NUM_LIMB, n);
    // both PAF channels of a limb with one 8-byte load: channel-last tensor, even channel count, aligned base
    const bool channel_last = paf.mode != PAF_PACKED && (paf.mode == PAF_FULL_HWC || paf.layout == EKP_LAYOUT_NHWC);
    const bool vec2 = channel_last && paf.C % 2 == 0 && reinterpret_cast<uintptr_t>(paf.ptr) % 8 == 0;
    // Blocks are latency-bound chains; a batch whose 19 x n blocks all fit on the GPU at once (crowded scenes come in
    // small batches) gets twice the threads per block, bigger batches keep more blocks resident instead.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool wide = EKP_NUM_LIMB * n <= 4 * sms;
#define EKP_LAUNCH_CONNECT(V, T) paf_connect_kernel<V, T><<<grid, T, 0, stream>>>(line, part_off, max_peaks, paf, h1, conns, n_conns, overflow)
    if (vec2) { if (wide) EKP_LAUNCH_CONNECT(true, 2 * kConnThreads); else EKP_LAUNCH_CONNECT(true, kConnThreads); }
    else { if (wide) EKP_LAUNCH_CONNECT(false, 2 * kConnThreads); else EKP_LAUNCH_CONNECT(false, kConnThreads); }
#undef EKP_LAUNCH_CONNECT
    return cudaGetLastError();
}

}  // namespace ekp
