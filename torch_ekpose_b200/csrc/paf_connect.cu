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
//  * sorting (pafprocess.cpp:97): the reference's result depends on HOW std::sort permutes equal
//    scores.  For n <= 16 libstdc++ runs a stable insertion sort, and without ties the order is
//    unique, so all threads rank the candidates in parallel (stable) and look for ties; only when
//    n > 16 AND ties exist does one thread replay libstdc++'s algorithm on the original sequence
//    (introsort: median-of-3 quicksort above 16 elements, heapsort after 2*floor(log2 n) levels,
//    then the final insertion sort), which reproduces the reference's permutation exactly.
#include "common.cuh"

namespace ekp {

constexpr int kConnThreads = 128;

struct Sample2 { float x, y; };

__device__ __forceinline__ Sample2 paf_sample(const PafSource& s, int img, int ly, int lx, int ch1, int ch2) {
    Sample2 r;
    lx = min(max(lx, 0), s.W - 1);  // memory safety only: valid peaks never sample outside
    ly = min(max(ly, 0), s.H - 1);
    if (s.mode == PAF_FULL_HWC) {
        const float* q = s.ptr + (((size_t) img * s.H + ly) * s.W + lx) * s.C;
        r.x = __ldg(q + ch1);
        r.y = __ldg(q + ch2);
    } else if (s.mode == PAF_LO_NEAREST) {
        r.x = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, ly >> 3, lx >> 3);
        r.y = lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, ly >> 3, lx >> 3);
    } else {  // identical arithmetic to materialise_tile (dense_frontend.cu)
        int i0, i1, j0, j1;
        float tx, ty;
        bilin_coord(lx, s.w, i0, i1, tx);
        bilin_coord(ly, s.h, j0, j1, ty);
        float top = lerp1(lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j0, i0), lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j0, i1), tx);
        float bot = lerp1(lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j1, i0), lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch1, j1, i1), tx);
        r.x = lerp1(top, bot, ty);
        top = lerp1(lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j0, i0), lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j0, i1), tx);
        bot = lerp1(lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j1, i0), lo_at(s.ptr, s.layout, img, s.C, s.h, s.w, ch2, j1, i1), tx);
        r.y = lerp1(top, bot, ty);
    }
    return r;
}

// pafprocess.cpp:59-94 for one (a, b) pair.  Returns true when the pair becomes a candidate.
__device__ __forceinline__ bool score_pair(const ekp_peak& a, const ekp_peak& b, const PafSource& paf, int img, int ch1,
                                           int ch2, int h1, float& criterion2) {
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
    for (int i = 0; i < 10; i++) sv[i] = paf_sample(paf, img, ly[i], lx[i], ch1, ch2);  // independent gathers in flight
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
// Control flow, comparisons and element moves are exactly libstdc++'s (same order), but every scan
// ("advance while comp holds") inspects 32 elements per step with a ballot and every shift of the
// insertion sort moves its elements in parallel.  All 32 lanes call these with identical arguments.
__device__ __forceinline__ int lead_true(unsigned m) { return m == 0xffffffffu ? 32 : __ffs(~m) - 1; }

// __unguarded_linear_insert(i) / the guarded branch of __insertion_sort: move element i left past
// every element it is greater than (`lower` = lowest index that may be inspected).
__device__ void warp_linear_insert(const CandArray& A, int i, int lower) {
    const int lane = threadIdx.x & 31;
    const Cand val = A.get(i);
    int d = 0;  // number of elements val has to pass
    for (;;) {
        const int idx = i - 1 - d - lane;
        const int run = lead_true(__ballot_sync(0xffffffffu, idx >= lower && val.s > A.s[idx]));
        d += run;
        if (run < 32) break;
    }
    if (d == 0) return;
    for (int c0 = 0; c0 < d; c0 += 32) {  // shift [i-d, i-1] up by one, topmost chunk first
        const int k = c0 + lane;
        Cand t;
        if (k < d) t = A.get(i - 1 - k);
        __syncwarp();
        if (k < d) A.set(i - k, t);
        __syncwarp();
    }
    if (lane == 0) A.set(i - d, val);
    __syncwarp();
}

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

__device__ void std_sort_desc(const CandArray& A, int n) {
    if (n <= 0) return;
    const int lane = threadIdx.x & 31;
    // __introsort_loop with an explicit stack: the recursion only ever touches disjoint ranges,
    // so the order in which they are finished does not change the result.
    int stk_first[48], stk_last[48], stk_depth[48];
    int sp = 0;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) lg++;
    stk_first[sp] = 0; stk_last[sp] = n; stk_depth[sp] = 2 * lg; sp++;
    while (sp) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {  // heapsort fallback: never reached by real scenes, kept serial
                if (lane == 0) sort_heapsort(A, first, last);
                __syncwarp();
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
            const int cut = warp_partition(A, first + 1, last, A.s[first], n);
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    // __final_insertion_sort: __insertion_sort on the first 16 (or all), then unguarded inserts; both
    // are "move left past everything smaller", bounded below by index 0
    for (int i = 1; i < n; ++i) warp_linear_insert(A, i, 0);
}

__global__ void __launch_bounds__(kConnThreads) paf_connect_kernel(const ekp_peak* __restrict__ line,
                                                                   const int* __restrict__ part_off, int max_peaks,
                                                                   const PafSource paf, int h1, Conn* __restrict__ conns,
                                                                   int* __restrict__ n_conns,
                                                                   unsigned* __restrict__ overflow) {
    __shared__ ekp_peak sA[EKP_MAX_PART], sB[EKP_MAX_PART];
    __shared__ float sScore[EKP_MAX_CAND];
    __shared__ unsigned sTag[EKP_MAX_CAND];
    __shared__ float sScore2[EKP_MAX_CAND];
    __shared__ unsigned sTag2[EKP_MAX_CAND];
    __shared__ int sTies;
    __shared__ int sWarpCnt[kConnThreads / 32];
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
    const ekp_peak* L = line + (size_t) img * max_peaks;
    for (int i = threadIdx.x; i < nA; i += kConnThreads) sA[i] = L[offA + i];
    for (int i = threadIdx.x; i < nB; i += kConnThreads) sB[i] = L[offB + i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npairs = nA * nB;
    int total = 0;  // candidates so far, identical in every thread
    for (int base = 0; base < npairs; base += kConnThreads) {
        const int pidx = base + threadIdx.x;
        bool pass = false;
        float crit = 0.f;
        int ia = 0, ib = 0;
        if (pidx < npairs) {
            ia = pidx / nB;
            ib = pidx - ia * nB;
            pass = score_pair(sA[ia], sB[ib], paf, img, ch1, ch2, h1, crit);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) sWarpCnt[warp] = __popc(mask);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int k = 0; k < kConnThreads / 32; k++) {
            const int c = sWarpCnt[k];
            if (k < warp) before += c;
            all += c;
        }
        if (pass) {
            const int pos = total + before + __popc(mask & ((1u << lane) - 1u));
            if (pos < EKP_MAX_CAND) { sScore[pos] = crit; sTag[pos] = ((unsigned) ia << 16) | (unsigned) ib; }
        }
        total += all;
        __syncthreads();
    }

    // ---- sort (pafprocess.cpp:97).  std::sort's result is only algorithm-dependent in how it
    // permutes EQUAL scores, and for n <= 16 it is a plain (stable) insertion sort.  So: rank every
    // candidate in parallel (stable order) and detect ties; only when n > 16 AND ties exist does
    // one thread replay libstdc++'s introsort on the original sequence.
    const int n = min(total, EKP_MAX_CAND);
    if (threadIdx.x == 0) {
        sTies = 0;
        if (total > EKP_MAX_CAND) atomicOr(overflow + img, EKP_OVF_CANDIDATES);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kConnThreads) {
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

    const bool replay = n > 16 && sTies;  // uniform
    if (replay && threadIdx.x < 32) {    // warp 0 replays libstdc++'s std::sort on the original sequence
        CandArray A;
        A.s = sScore; A.t = sTag;
        std_sort_desc(A, n);
    }
    if (threadIdx.x == 0) {
        const float* srcS = replay ? sScore : sScore2;
        const unsigned* srcT = replay ? sTag : sTag2;
        // greedy assignment, pafprocess.cpp:98-124 (a peak is used at most once per limb side)
        unsigned usedA[EKP_MAX_PART / 32], usedB[EKP_MAX_PART / 32];
#pragma unroll
        for (int k = 0; k < EKP_MAX_PART / 32; k++) usedA[k] = usedB[k] = 0u;
        Conn* out = conns + ((size_t) img * EKP_NUM_LIMB + limb) * EKP_MAX_PART;
        int nc = 0;
        for (int c = 0; c < n; c++) {
            const unsigned tag = srcT[c];
            const int i1 = tag >> 16, i2 = tag & 0xffff;
            if ((usedA[i1 >> 5] >> (i1 & 31)) & 1u) continue;
            if ((usedB[i2 >> 5] >> (i2 & 31)) & 1u) continue;
            usedA[i1 >> 5] |= 1u << (i1 & 31);
            usedB[i2 >> 5] |= 1u << (i2 & 31);
            Conn cn;
            cn.cid1 = sA[i1].id; cn.cid2 = sB[i2].id; cn.score = srcS[c]; cn.pad = 0;
            out[nc++] = cn;
        }
        *out_n = nc;
    }
}

// Test hook: the device replay of libstdc++'s std::sort on caller-supplied scores (one warp), so that the tie permutation -- including the heapsort fallback, which real scenes never
// reach -- can be compared with the compiled reference's std::sort.
__global__ void debug_std_sort_kernel(float* scores, unsigned* tags, int n) {
    CandArray A;  // one warp, as in paf_connect_kernel
    A.s = scores; A.t = tags;
    std_sort_desc(A, n);
}
cudaError_t launch_debug_std_sort(float* scores, unsigned* tags, int n, cudaStream_t stream) {
    debug_std_sort_kernel<<<1, 32, 0, stream>>>(scores, tags, n);
    return cudaGetLastError();
}

cudaError_t launch_paf_connect(const ekp_peak* line, const int* part_off, int max_peaks, const PafSource& paf, int h1,
                               int n, Conn* conns, int* n_conns, unsigned* overflow, cudaStream_t stream) {
    dim3 grid(EKP_NUM_LIMB, n);
    paf_connect_kernel<<<grid, kConnThreads, 0, stream>>>(line, part_off, max_peaks, paf, h1, conns, n_conns, overflow);
    return cudaGetLastError();
}

}  // namespace ekp
