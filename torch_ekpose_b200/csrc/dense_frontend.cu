// dense_frontend.cu -- stages 1-3 of the hot path, fused, for sm_100a.
//
// One launch takes the network's stride-8 heat (19 ch) and PAF (38 ch) maps of a whole batch
// and, per 16-row x (8*tile_wl)-column full-resolution tile (one CTA),
//   (0) first wave of CTAs only: prefetches the whole batch's inputs into L2 (evict_last, TMA bulk
//       prefetch) in one burst, so input reads do not trickle in between the writes,
//   (1) stages the stride-8 neighbourhood of the tile in shared memory (HWC order) with cp.async,
//   (2) optionally materialises the bilinear x8 tensors heat_mat[H][W][19] / paf_mat[H][W][38]
//       (the operator-surface tensors of process_paf, paf_to_pose.py:356-360) -- the HBM-bound part:
//       36.25 MB written per 368x432 image against 0.57 MB read.  The rows are built in
//       shared-memory store buffers and written by the TMA engine (cp.async.bulk shared -> global,
//       4 KB contiguous per copy), so no warp ever waits on a store,
//   (3) evaluates the Gaussian-smoothed (sigma 3, 25 taps, reflect) bilinear-upsampled heat map
//       as ONE separable 5-tap polyphase filter on the stride-8 grid (the composition of the two
//       linear operators; tables built on the host in capi.cu) entirely in registers -- only for
//       the (part, 30-pixel strip) pairs that can contain a value above the threshold at all
//       (exact early-out from a per-tile column-maximum table),
//   (4) does the 3x3 max NMS with warp shuffles (x) and a rolling 3-row window (y) and appends
//       peaks with a warp-ballot aggregated atomic into the per-image raw peak list.
// When materialising, the CTA's warps are specialised: "fill" warps feed the TMA engine, the others do
// (3)-(4) underneath the store stream (process_tile_mat).  Nothing but the (optional)
// operator-surface tensors and the peaks ever goes back to HBM.
//
// Arithmetic is the one defined in oracle/frontend_oracle.c part (B); results are bit-identical
// to it (tests/test_gpu_parity.py).  There is no reference implementation of this front-end
// (SURVEY.md 0.1); the reference's own front-end is ref_frontend.cu.
//
// TMA: the OUTPUT side uses the TMA engine's plain bulk copies (16-byte aligned row segments).  The
// INPUT side cannot: cuTensorMap strides must be multiples of 16 bytes, the rows of these tensors
// are 54 (or 82, 164) floats and the pixels 19 / 38 floats, so neither the NCHW nor the NHWC input
// can be described by a tiled tensor map without re-padding it, and the NCHW -> HWC transpose
// that the staging performs on the fly is not a box copy.  4-byte cp.async (LDGSTS) does both.
//
// History (profiles/README.md): issue-slot bound at 0.50 of the measured HBM roofline (666 M warp
// instructions) -> per-thread streaming stores at 0.92 (145 M) -> TMA bulk stores + warp roles 0.94.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace ekp {

#ifndef EKP_LEAN_THREADS
#define EKP_LEAN_THREADS 256
#endif
#ifndef EKP_LEAN_BLOCKS
#define EKP_LEAN_BLOCKS 5
#endif
#ifndef EKP_MAT_THREADS
#define EKP_MAT_THREADS 384
#endif
#ifndef EKP_MAT_BLOCKS
#define EKP_MAT_BLOCKS 2
#endif
#ifndef EKP_FILL_WARPS
#define EKP_FILL_WARPS 8
#endif
#ifndef EKP_STORE_BUFS
#define EKP_STORE_BUFS 2
#endif
#ifndef EKP_CHUNK_COLS
#define EKP_CHUNK_COLS 256
#endif
#ifndef EKP_MAX_TWL
#define EKP_MAX_TWL 32
#endif

constexpr int kMaxTwl = EKP_MAX_TWL;          // widest tile in stride-8 columns
constexpr int kLeanThreads = EKP_LEAN_THREADS;  // CTA size / resident CTAs per SM without materialisation
constexpr int kLeanBlocks = EKP_LEAN_BLOCKS;
constexpr int kMatThreads = EKP_MAT_THREADS;    // ... with it (the store buffers take shared memory: 2 CTAs per SM)
constexpr int kMatBlocks = EKP_MAT_BLOCKS;
constexpr int kFillWarps = EKP_FILL_WARPS;      // warps that fill the store buffers and drive the TMA engine
constexpr int kFillThreads = 32 * kFillWarps;
constexpr int kNmsThreads = kMatThreads - kFillThreads;  // the other warps: heat patch, early-out, smoothing + NMS
constexpr int kStoreBufs = EKP_STORE_BUFS;      // store buffers of [kChunkRows][kChunkCols float4] per CTA
constexpr int kChunkCols = EKP_CHUNK_COLS;      // float4 columns per store chunk: one bulk copy = kChunkCols * 16 B of a row
constexpr int kColsPerFillThread = kChunkCols / kFillThreads;
#ifndef EKP_CHUNK_ROWS
#define EKP_CHUNK_ROWS 8
#endif
constexpr int kChunkRows = EKP_CHUNK_ROWS;      // output rows per store chunk (8 = a whole stride-8 row pair, or 4)
constexpr int kStoreBufF4 = kChunkRows * kChunkCols;
static_assert(kChunkRows == 8 || kChunkRows == 4, "a chunk is a whole or half row pair");
static_assert(kFillWarps >= 1 && kNmsThreads >= 32 && kNmsThreads % 32 == 0 && kChunkCols % kFillThreads == 0, "warp roles");
enum { BAR_FILL = 1, BAR_NMS = 2, BAR_HEAT_READY = 3 };  // named barriers (0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#ifndef EKP_LEAN_TB
#define EKP_LEAN_TB 4
#endif
constexpr int kTB = 2;                    // stride-8 row blocks per tile of the materialising kernel (16 output rows)
constexpr int kLeanTB = EKP_LEAN_TB;      // ... of the lean kernel for big batches: the 6-row halo of the window weighs less
                                          // (small batches keep kTB: more, lighter CTAs balance better)
constexpr int kTaskTB = 2;                // row blocks per NMS task: a tall tile is cut into sub-tiles of this height, so the
                                          // early-out stays as selective (and the tasks as balanced) as with 16-row tiles
constexpr int kMaxSubTiles = ((kLeanTB > kTB ? kLeanTB : kTB) + kTaskTB - 1) / kTaskTB;
constexpr int kPafRows = kTB + 1;         // rows staged for the tile's bilinear row pairs (m0-1 .. m0+tb-1)

// Vertical taps of an interior row (no reflect / clamp influence) depend only on Y & 7.
__constant__ float cTapsInterior[8][8];

// cTapsInteriorMax[j] = max over the 8 phases of cTapsInterior[.][j]: the most row j of the window can weigh
__constant__ float cTapsInteriorMax[8];

cudaError_t set_interior_taps(const float* taps64) {
    float mx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 8; k++)
        for (int j = 0; j < 8; j++) mx[j] = taps64[k * 8 + j] > mx[j] ? taps64[k * 8 + j] : mx[j];
    cudaError_t e = cudaMemcpyToSymbol(cTapsInterior, taps64, sizeof(float) * 64);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(cTapsInteriorMax, mx, sizeof(mx));
    return e;
}

// floor(n / d) == __umulhi(n, magic_of(d)) for n * d < 2^32 (32-bit divide only: no 64-bit division subroutine)
__host__ __device__ __forceinline__ unsigned magic_of(unsigned d) { return 0xFFFFFFFFu / d + 1u; }
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned magic) { return __umulhi(n, magic); }  // n, d < 2^16

// ---- (1) stage a stride-8 patch in shared memory as [row][col][C] ----------------------------
// Asynchronous 4-byte copies (cp.async / LDGSTS): a thread issues ALL of its ~37 element copies
// back to back without holding registers, so the whole patch costs one global-memory round trip
// instead of one per unrolled batch of loads.  The caller waits with stage_wait().
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

template <int C>
__device__ __forceinline__ void stage_patch(float* __restrict__ sP, const float* __restrict__ src, int layout,
                                            int img, int h, int w, int r0, int r1, int c0, int c1, int pcols,
                                            unsigned tid, unsigned nthr) {
    const int nr = r1 - r0 + 1, nc = c1 - c0 + 1;
    if (layout == EKP_LAYOUT_NCHW) {
        const unsigned plane = nr * nc, total = C * plane;
        const unsigned m_plane = magic_of(plane), m_nc = magic_of(nc);  // uniform, a handful of instructions
        const float* g0 = src + ((size_t) img * C * h + r0) * w + c0;
        const int hw = h * w;
        for (unsigned idx = tid; idx < total; idx += nthr) {
            const unsigned c = fastdiv(idx, m_plane);
            const unsigned rem = idx - c * plane;
            const unsigned r = fastdiv(rem, m_nc);
            const unsigned i = rem - r * nc;
            cp_async_f32(sP + (r * pcols + i) * C + c, g0 + c * hw + r * w + i);
        }
    } else {
        const unsigned per_r = nc * C, total = nr * per_r;
        const unsigned m_row = magic_of(per_r);
        const float* g0 = src + (((size_t) img * h + r0) * w + c0) * C;
        for (unsigned idx = tid; idx < total; idx += nthr) {
            const unsigned r = fastdiv(idx, m_row);
            const unsigned rem = idx - r * per_r;
            cp_async_f32(sP + r * pcols * C + rem, g0 + (size_t) r * w * C + rem);
        }
    }
}

// ---- (2) materialise heat_mat / paf_mat through the TMA engine --------------------------------
// cp.async.bulk (UBLKCP) shared -> global: a chunk of [<= 8 output rows][kChunkCols float4 columns]
// is written to a shared-memory buffer by the fill warps (a thread owns a float4 column = 4
// consecutive (x, c) entries of an output row: the horizontal interpolation is done once per
// stride-8 row, then each output row costs four FMAs and one 16-byte shared-memory store), then one
// lane per fill warp hands one row each (kChunkCols x 16 B = 4 KB contiguous in HBM) to the TMA
// engine.  Stores never occupy the warps' scoreboards or LSU queues: the engine drains the buffers
// while the CTA stages, smooths and does the NMS.  tools/bulk_store_probe.cu: this pattern alone
// sustains 7.0-7.2 TB/s against 6.6-6.9 TB/s for per-thread st.global.cs.v4 on the same tiles.
// A plain (non-tensor) bulk copy only needs 16-byte aligned addresses and sizes, which every row
// segment of heat_mat / paf_mat has (W = 8w, so a row is 32*w*C bytes and a tile starts at
// 32*i0*C bytes); no tensor map is involved.
// The L2 policy of the stores matters: evict_first lets L2 write these lines back early and in order instead of
// ageing them through the LRU with everything else (they are never read again by this kernel): 168.8 k -> 174.8 k
// images/s; evict_unchanged / no hint are equal (profiles/README.md).
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_store_row(void* gdst, const void* ssrc, unsigned bytes, unsigned long long policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"((unsigned) __cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Called by the kFillThreads threads of the fill warps (named barrier BAR_FILL inside); t is the thread's
// index within that group.  A thread owns kColsPerFillThread float4 columns of the chunk.  `phase`
// counts the chunks emitted so far by this CTA (buffer ring position; lane 0 of every fill warp
// commits one bulk group per chunk, so "all but the kStoreBufs-1 most recent groups have been read"
// means the next buffer is free).
template <int C>
__device__ __forceinline__ void materialise_tile(const float* __restrict__ sP, int pr0, int pc0, int pcols,
                                                      float* __restrict__ out_img, int h, int w, int m0, int tb, int i0,
                                                      int twl, float4* __restrict__ sStore, int& phase, int t) {
    constexpr int S = kColsPerFillThread;
    const int W = w * 8;
    const int X0 = i0 * 8;
    const int row_f4 = twl * 2 * C;
    const size_t stride4 = (size_t) W * C / 4;
    const int prow = pcols * C;
    const int lane = t & 31, fwarp = t >> 5;
    const int Ystart = max(8 * m0 - 4, 0);
    const unsigned long long store_policy = l2_policy_evict_first();
    for (int c0 = 0; c0 < row_f4; c0 += kChunkCols) {
        const int ncol = min(kChunkCols, row_f4 - c0);
        bool valid[S];
        int off0[S][4], off1[S][4];
        float tx[S][4];
#pragma unroll
        for (int s = 0; s < S; s++) {
            valid[s] = s * kFillThreads + t < ncol;
            const int col = c0 + (valid[s] ? s * kFillThreads + t : 0);  // idle slots shadow column c0 (never stored)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int f = col * 4 + e;
                const int xl = f / C;
                const int c = f - xl * C;
                int a, b;
                bilin_coord(X0 + xl, w, a, b, tx[s][e]);
                off0[s][e] = (a - pc0) * C + c;
                off1[s][e] = (b - pc0) * C + c;
            }
        }
        float top[S][4], bot[S][4], d[S][4];
        auto load_row = [&](int j) {  // horizontal lerp of stride-8 row j
            const float* r = sP + (j - pr0) * prow;
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    top[s][e] = bot[s][e];
                    bot[s][e] = lerp1(r[off0[s][e]], r[off1[s][e]], tx[s][e]);
                    d[s][e] = __fsub_rn(bot[s][e], top[s][e]);
                }
        };
        float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) Ystart * W + X0) * C) + c0;
        auto emit_chunk = [&](auto k0_tag, auto n_tag) {  // rows k0 .. k0+n-1 of the current stride-8 row pair
            constexpr int K0 = decltype(k0_tag)::value, NN = decltype(n_tag)::value;
            float4* buf = sStore + (size_t) (phase % kStoreBufs) * kStoreBufF4;
            if (lane == 0) bulk_wait_read<kStoreBufs - 1>();  // the copies that last read this buffer are done reading
            bar_sync(BAR_FILL, kFillThreads);
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (!valid[s]) continue;
#pragma unroll
                for (int k = 0; k < NN; k++) {
                    const float ty = (float) (2 * (K0 + k) + 1) * 0.0625f;
                    float4 v;
                    v.x = fmaf(ty, d[s][0], top[s][0]);
                    v.y = fmaf(ty, d[s][1], top[s][1]);
                    v.z = fmaf(ty, d[s][2], top[s][2]);
                    v.w = fmaf(ty, d[s][3], top[s][3]);
                    buf[k * kChunkCols + s * kFillThreads + t] = v;
                }
            }
            fence_proxy_async();  // generic-proxy writes above -> visible to the async proxy (TMA)
            bar_sync(BAR_FILL, kFillThreads);
            if (lane == 0) {
                for (int k = fwarp; k < NN; k += kFillWarps) bulk_store_row(dst + (size_t) k * stride4, buf + k * kChunkCols, ncol * 16, store_policy);
                bulk_commit();
            }
            dst += (size_t) NN * stride4;
            phase++;
        };
        auto emit = [&](auto k0_tag, auto n_tag) {  // ... in chunks of kChunkRows rows
            constexpr int K0 = decltype(k0_tag)::value, NN = decltype(n_tag)::value;
            if constexpr (NN > kChunkRows) {
                emit_chunk(std::integral_constant<int, K0>{}, std::integral_constant<int, kChunkRows>{});
                emit_chunk(std::integral_constant<int, K0 + kChunkRows>{}, std::integral_constant<int, NN - kChunkRows>{});
            } else {
                emit_chunk(k0_tag, n_tag);
            }
        };
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
            for (int e = 0; e < 4; e++) bot[s][e] = 0.f;
        using I0 = std::integral_constant<int, 0>;
        using I4 = std::integral_constant<int, 4>;
        using I8 = std::integral_constant<int, 8>;
        // The tile materialises the rows of the stride-8 row PAIRS (m0-1, m0) .. (m0+tb-2, m0+tb-1),
        // i.e. rows 8*m0-4 .. 8*(m0+tb)-5: each pair is complete (8 rows, one horizontal lerp per
        // stride-8 row).  The clamped half pairs at the top and bottom of the image belong to the
        // first and last tile.
        load_row(max(m0 - 1, 0));
        load_row(m0);
        if (m0 == 0) emit(I4{}, I4{});
        else emit(I0{}, I8{});
        for (int q = 1; q < tb; q++) {
            load_row(m0 + q);
            emit(I0{}, I8{});
        }
        if (m0 + tb == h) {
            load_row(h - 1);
            emit(I0{}, I4{});
        }
    }
}

// ---- (3)+(4) smoothed map and NMS ------------------------------------------------------------
struct NmsState {
    float hm2, hm1;  // horizontal 3-max of rows Y-2 and Y-1
    float s1;        // smoothed value of row Y-1 (-inf when that row is not owned by the tile)
};

struct PeakSink {
    RawPeak* raw;
    int* count;
    int cap;
};

__device__ __noinline__ void emit_peaks(const PeakSink& sink, bool is_peak, int X, int Y, float score, int part) {
    const unsigned mask = __ballot_sync(0xffffffffu, is_peak);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(sink.count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (is_peak) {
        const int slot = base + __popc(mask & ((1u << lane) - 1u));
        if (slot < sink.cap) {
            RawPeak pk;
            pk.x = X; pk.y = Y; pk.score = score; pk.part = part;
            pk.key = ((unsigned) Y << 16) | (unsigned) X;
            sink.raw[slot] = pk;
        }
    }
}

// One full-resolution row: S is this lane's smoothed value at row Y (already -inf outside the
// image).  Tests row Y-1, then rolls the state.
__device__ __forceinline__ void nms_row(NmsState& st, float S, int X, int Y, bool out_lane, float thr, int part,
                                        const PeakSink& sink) {
    const float l = __shfl_up_sync(0xffffffffu, S, 1);
    const float r = __shfl_down_sync(0xffffffffu, S, 1);
    const float hm = fmaxf(S, fmaxf(l, r));
    const bool is_peak = out_lane && st.s1 > thr && st.s1 == fmaxf(st.hm2, fmaxf(st.hm1, hm));
    if (__any_sync(0xffffffffu, is_peak)) emit_peaks(sink, is_peak, X, Y - 1, st.s1, part);
    st.hm2 = st.hm1;
    st.hm1 = hm;
    st.s1 = S;
}

// ---- one tile ------------------------------------------------------------------------------------
struct TileGeom {
    int img, m0, i0, twl, tb;
    int hr0, hr1, hc0, hc1;  // staged heat rows / columns (smoothing window +-3 cells, clamped like the taps)
    int pr0, pr1, pc0, pc1;  // staged PAF rows / columns (bilinear row pairs of the tile)
};

template <int TB>
__device__ __forceinline__ TileGeom tile_geom(const DenseParams& p, int tile_x, int tile_y, int img) {
    TileGeom g;
    const int h = p.h, w = p.w;
    g.img = img;
    g.m0 = tile_y * TB;
    g.i0 = tile_x * p.tile_wl;
    g.twl = min(p.tile_wl, w - g.i0);
    g.tb = min(TB, h - g.m0);
    // the smoothing window of row block m is rows clamp(m-2, 0, h-5) .. +4 (same for columns)
    g.hr0 = max(min(g.m0 - 3, h - 5), 0); g.hr1 = min(max(g.m0 + g.tb + 2, 4), h - 1);
    g.hc0 = max(min(g.i0 - 3, w - 5), 0); g.hc1 = min(max(g.i0 + g.twl + 2, 4), w - 1);
    g.pr0 = max(g.m0 - 1, 0); g.pr1 = min(g.m0 + g.tb - 1, h - 1);
    g.pc0 = max(g.i0 - 1, 0); g.pc1 = min(g.i0 + g.twl, w - 1);
    return g;
}

struct TileSmem {
    float* heat;    // [tile blocks + 6][tile_wl + 6][19]
    float* paf;     // [kPafRows][tile_wl + 2][38]
    float4* store;  // [kStoreBufs][kChunkRows][kChunkCols] float4 (materialising kernel only)
};

// Everything a tile's threads share besides the patches.
struct TileCtl {
    float* colmax;            // [tile_wl + 6][19] max(0, column maximum over the staged heat rows)
    const float* taps;        // [8][8] interior taps by phase (copy of cTapsInterior for per-lane indexing)
    unsigned short* list;     // (part, strip) tasks that survive the early-out
    int* num_active;
    int* next_task;
};

// per (column, channel) maximum of the staged heat samples, for the early-out below
__device__ __forceinline__ void build_colmax(const TileGeom& g, const float* sHeat, int hcols, float* sColMax, int tid, int nthr) {
    const int ncolc = (g.hc1 - g.hc0 + 1) * EKP_HEAT_CH, nrow = g.hr1 - g.hr0 + 1;
    for (int idx = tid; idx < ncolc; idx += nthr) {
        float mx = 0.f;
        for (int r = 0; r < nrow; r++) mx = fmaxf(mx, sHeat[r * hcols * EKP_HEAT_CH + idx]);
        sColMax[idx] = mx;
    }
}

// ---- exact early-out, decided once per (part, strip) by one thread each -------------------------
// All taps are >= 0 and sum to 1 (checked on the host), so every smoothed value a strip can
// produce is a convex combination of the staged stride-8 samples in columns
// [bx(first lane), bx(last lane) + 4]: if their maximum (sColMax, over all staged rows) is below
// the threshold by more than the rounding slack of ten float operations, no pixel there can
// pass `S > thr`, hence no peak, and the strip is never visited.
// Strips that pass this cheap test get a sharper one, row block by row block: with M_j the maximum (>= 0) of
// window row j over the strip's columns, the horizontal pass gives T_j <= M_j for every lane, so a pixel of an
// interior row block is at most sum_j max_phase(tap_j) * M_j -- the outer rows of the window weigh a few
// percent, so a blob two cells above or below the tile no longer activates it (about half of the tasks).
// Border blocks (reflected taps) keep the plain maximum.  Survivors go to a compact list that the warps share.
template <bool kDebug>
__device__ __forceinline__ void build_task_list(const DenseParams& p, const TileGeom& g, const float* sHeat, const TileCtl& ctl,
                                                int tid, int nthr) {
    const int h = p.h, w = p.w, W = 8 * w, TW = 8 * g.twl, X0 = 8 * g.i0;
    const int hcols = p.tile_wl + 6;
    const int nstrips = (TW + 29) / 30;
    const unsigned m_strips = magic_of(nstrips);
    const int per_sub = EKP_NUM_PART * nstrips;
    const int ntask = per_sub * ((g.tb + kTaskTB - 1) / kTaskTB);  // task = (sub-tile, part, strip)
    for (int t = tid; t < ntask; t += nthr) {
        const int sub = t >= per_sub ? t / per_sub : 0, t_in = t - sub * per_sub;  // (no division in 16-row tiles)
        const int c = (int) fastdiv(t_in, m_strips);
        const int strip = t_in - c * nstrips;
        const int b_lo = sub * kTaskTB, b_hi = min(b_lo + kTaskTB, g.tb);  // this task's row blocks within the tile
        bool active = kDebug || !(p.thr > 0.f);
        if (!active) {
            const int Xa = X0 - 1 + 30 * strip;
            const int xlo = min(max(Xa, 0), min(W - 1, X0 + TW)), xhi = min(max(Xa + 31, 0), min(W - 1, X0 + TW));
            const int c_lo = min(max((xlo >> 3) - 2, 0), w - 5), c_hi = min(max((xhi >> 3) - 2, 0), w - 5) + 4;
            float mx = 0.f;
            for (int i = c_lo; i <= c_hi; i++) mx = fmaxf(mx, ctl.colmax[(i - g.hc0) * EKP_HEAT_CH + c]);
            active = mx > p.thr * 0.99999f;
            if (active) {
                bool any = false;
                for (int b = b_lo; b < b_hi && !any; b++) {
                    const int m = g.m0 + b;
                    const int wb = min(max(m - 2, 0), h - 5);
                    const bool interior = m >= 2 && m <= h - 3;
                    float bound = 0.f;
#pragma unroll
                    for (int j = 0; j < 5; j++) {
                        const float* row = sHeat + ((wb + j - g.hr0) * hcols - g.hc0) * EKP_HEAT_CH + c;
                        float mj = 0.f;
                        for (int i = c_lo; i <= c_hi; i++) mj = fmaxf(mj, row[i * EKP_HEAT_CH]);
                        bound = interior ? fmaf(cTapsInteriorMax[j], mj, bound) : fmaxf(bound, mj);
                    }
                    any = bound > p.thr * 0.9999f;
                }
                active = any;
            }
        }
        if (active) ctl.list[atomicAdd(ctl.num_active, 1)] = (unsigned short) t;
    }
}

// One (part, 30-column strip, <= kTaskTB row blocks) task, by one warp: smoothed map of the task's rows (+ one halo row
// above and below) in registers, 3x3 max NMS, peaks appended to the image's raw list.  The staged stride-8 samples of
// part c are at base[j * rstride + i * CS] (row j, column i of the whole map): CS = 19 for the HWC tile patch of the
// tiled kernels, 1 for the plane of the plane kernel.  X0 / TW: first full-resolution column / width the task's strips
// tile; m0 / tb: first stride-8 row block / number of row blocks.
template <bool kDebug, int CS>
__device__ __forceinline__ void nms_task_at(const DenseParams& p, int img, int c, int strip, int m0, int tb, int X0, int TW,
                                            const float* __restrict__ base, int rstride, const float* __restrict__ sTaps) {
    const int h = p.h, w = p.w, H = 8 * h, W = 8 * w;
    const int lane = threadIdx.x & 31;
    const float NEG_INF = __int_as_float(0xff800000);
    PeakSink sink;
    sink.raw = p.raw + (size_t) img * p.raw_cap;
    sink.count = p.raw_count + img;
    sink.cap = p.raw_cap;

    const int X = X0 - 1 + 30 * strip + lane;
    const bool inb = X >= 0 && X < W;
    const bool out_lane = lane >= 1 && lane <= 30 && X < X0 + TW && X < W;
    const int Xc = min(max(X, 0), min(W - 1, X0 + TW));  // lanes past the halo are never outputs
    const int bx = min(max((Xc >> 3) - 2, 0), w - 5);
    const float* colp = base + bx * CS;

    // horizontal taps of this lane's column (only now: a skipped strip must not pay for the loads);
    // interior columns take them from the phase table in shared memory, border columns from global
    float4 axv;
    float ax4;
    if ((Xc >> 3) >= 2 && (Xc >> 3) <= w - 3) {
        axv = *reinterpret_cast<const float4*>(sTaps + (Xc & 7) * 8);
        ax4 = sTaps[(Xc & 7) * 8 + 4];
    } else {
        axv = __ldg(reinterpret_cast<const float4*>(p.ax + (size_t) Xc * 8));
        ax4 = __ldg(p.ax + (size_t) Xc * 8 + 4);
    }
    auto trow = [&](int j) -> float {  // horizontal 5-tap pass on stride-8 row j
        const float* s = colp + j * rstride;
        float acc = __fmul_rn(axv.x, s[0]);
        acc = fmaf(axv.y, s[CS], acc);
        acc = fmaf(axv.z, s[2 * CS], acc);
        acc = fmaf(axv.w, s[3 * CS], acc);
        acc = fmaf(ax4, s[4 * CS], acc);
        return acc;
    };
    float T0 = 0.f, T1 = 0.f, T2 = 0.f, T3 = 0.f, T4 = 0.f;
    int cur_wb = -100;
    auto window = [&](int m) {  // make T0..T4 the horizontal results of row block m's window
        const int wb = min(max(m - 2, 0), h - 5);
        if (wb == cur_wb) return;
        if (wb == cur_wb + 1) { T0 = T1; T1 = T2; T2 = T3; T3 = T4; T4 = trow(wb + 4); }
        else { T0 = trow(wb); T1 = trow(wb + 1); T2 = trow(wb + 2); T3 = trow(wb + 3); T4 = trow(wb + 4); }
        cur_wb = wb;
    };
    auto row_generic = [&](int Y) -> float {  // border rows: taps from the table in global memory
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.ay + (size_t) Y * 8));
        const float a4 = __ldg(p.ay + (size_t) Y * 8 + 4);
        float acc = __fmul_rn(a.x, T0);
        acc = fmaf(a.y, T1, acc);
        acc = fmaf(a.z, T2, acc);
        acc = fmaf(a.w, T3, acc);
        acc = fmaf(a4, T4, acc);
        return acc;
    };
    auto debug_out = [&](int Y, float acc) {
        if (kDebug && out_lane) p.smooth_out[(((size_t) img * H + Y) * W + X) * EKP_NUM_PART + c] = acc;
    };

    NmsState st;
    st.hm2 = NEG_INF; st.hm1 = NEG_INF; st.s1 = NEG_INF;
    const int Ytop = 8 * m0 - 1;
    if (!kDebug && m0 >= 3 && m0 + tb <= h - 3) {
        // Interior task (most of a map): every window, the halo rows' included, is row block m's own five rows m - 2 .. m + 2
        // and the next one is the previous one shifted by a row; the halo rows 8 m0 - 1 and 8 (m0 + tb) are interior rows too,
        // so their taps are phases 7 and 0 of the constant-bank table (ensure_tables checks that the table in global memory
        // repeats it there).  Same operations as the general walk below without its window bookkeeping and global loads.
        auto interior_row = [&](int k) -> float {
            float acc = __fmul_rn(cTapsInterior[k][0], T0);
            acc = fmaf(cTapsInterior[k][1], T1, acc);
            acc = fmaf(cTapsInterior[k][2], T2, acc);
            acc = fmaf(cTapsInterior[k][3], T3, acc);
            acc = fmaf(cTapsInterior[k][4], T4, acc);
            return acc;
        };
        T0 = trow(m0 - 3); T1 = trow(m0 - 2); T2 = trow(m0 - 1); T3 = trow(m0); T4 = trow(m0 + 1);
        {
            const float acc = interior_row(7);
            nms_row(st, inb ? acc : NEG_INF, X, Ytop, false, p.thr, c, sink);
            st.s1 = NEG_INF;
        }
        for (int b = 0; b < tb; b++) {
            const int m = m0 + b;
            T0 = T1; T1 = T2; T2 = T3; T3 = T4; T4 = trow(m + 2);
            float bound = __fmul_rn(cTapsInteriorMax[0], fmaxf(T0, 0.f));
            bound = fmaf(cTapsInteriorMax[1], fmaxf(T1, 0.f), bound);
            bound = fmaf(cTapsInteriorMax[2], fmaxf(T2, 0.f), bound);
            bound = fmaf(cTapsInteriorMax[3], fmaxf(T3, 0.f), bound);
            bound = fmaf(cTapsInteriorMax[4], fmaxf(T4, 0.f), bound);
            if (!__any_sync(0xffffffffu, inb && bound > p.thr * 0.9999f)) {
                nms_row(st, NEG_INF, X, 8 * m, out_lane, p.thr, c, sink);
                st.hm2 = NEG_INF; st.hm1 = NEG_INF; st.s1 = NEG_INF;
                continue;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float acc = interior_row(k);
                nms_row(st, inb ? acc : NEG_INF, X, 8 * m + k, out_lane, p.thr, c, sink);
            }
        }
        T0 = T1; T1 = T2; T2 = T3; T3 = T4; T4 = trow(m0 + tb + 2);
        const float accb = interior_row(0);
        nms_row(st, inb ? accb : NEG_INF, X, 8 * (m0 + tb), out_lane, p.thr, c, sink);
        return;
    }
    if (Ytop >= 0) {  // halo row above the task's rows: contributes its horizontal max only
        window(m0 - 1);
        const float acc = row_generic(Ytop);
        nms_row(st, inb ? acc : NEG_INF, X, Ytop, false, p.thr, c, sink);
        st.s1 = NEG_INF;  // not an owned row: never reported from here
    }
    for (int b = 0; b < tb; b++) {
        const int m = m0 + b;
        window(m);
        if (m >= 2 && m <= h - 3) {  // interior block: taps are immediates from the constant bank
            // Exact early-out per row block from the lanes' own horizontal results: taps >= 0, so no row of the block
            // exceeds sum_j max_phase(tap_j) * max(T_j, 0) in this lane's column; if that stays below the threshold in every
            // lane, none of the 8 rows holds a peak, and as neighbours they cannot outweigh a value above the threshold
            // either: the pending row is tested against -inf and the block is skipped (same peaks, tests/test_gpu_parity.py).
            float bound = __fmul_rn(cTapsInteriorMax[0], fmaxf(T0, 0.f));
            bound = fmaf(cTapsInteriorMax[1], fmaxf(T1, 0.f), bound);
            bound = fmaf(cTapsInteriorMax[2], fmaxf(T2, 0.f), bound);
            bound = fmaf(cTapsInteriorMax[3], fmaxf(T3, 0.f), bound);
            bound = fmaf(cTapsInteriorMax[4], fmaxf(T4, 0.f), bound);
            if (!kDebug && !__any_sync(0xffffffffu, inb && bound > p.thr * 0.9999f)) {
                nms_row(st, NEG_INF, X, 8 * m, out_lane, p.thr, c, sink);   // tests row 8m - 1
                st.hm2 = NEG_INF; st.hm1 = NEG_INF; st.s1 = NEG_INF;        // rows 8m .. 8m + 7: nothing above the threshold
                continue;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float acc = __fmul_rn(cTapsInterior[k][0], T0);
                acc = fmaf(cTapsInterior[k][1], T1, acc);
                acc = fmaf(cTapsInterior[k][2], T2, acc);
                acc = fmaf(cTapsInterior[k][3], T3, acc);
                acc = fmaf(cTapsInterior[k][4], T4, acc);
                debug_out(8 * m + k, acc);
                nms_row(st, inb ? acc : NEG_INF, X, 8 * m + k, out_lane, p.thr, c, sink);
            }
        } else {
#pragma unroll 2
            for (int k = 0; k < 8; k++) {
                const float acc = row_generic(8 * m + k);
                debug_out(8 * m + k, acc);
                nms_row(st, inb ? acc : NEG_INF, X, 8 * m + k, out_lane, p.thr, c, sink);
            }
        }
    }
    const int Ybot = 8 * (m0 + tb);  // halo row below (or the virtual row below the image)
    float sb = NEG_INF;
    if (Ybot < H) {
        window(m0 + tb);
        const float acc = row_generic(Ybot);
        if (inb) sb = acc;
    }
    nms_row(st, sb, X, Ybot, out_lane, p.thr, c, sink);
}

// ... of a tile of the tiled kernels: task = (sub-tile, part, strip), samples in the tile's HWC patch
template <bool kDebug>
__device__ __forceinline__ void nms_task(const DenseParams& p, const TileGeom& g, const float* sHeat, const float* sTaps, int task) {
    const int hcols = p.tile_wl + 6;
    const int TW = 8 * g.twl, X0 = 8 * g.i0;
    const int nstrips = (TW + 29) / 30;
    const int sub = task >= EKP_NUM_PART * nstrips ? task / (EKP_NUM_PART * nstrips) : 0;
    task -= sub * (EKP_NUM_PART * nstrips);
    const int m0 = g.m0 + sub * kTaskTB, tb = min(kTaskTB, g.m0 + g.tb - m0);  // rows of this sub-tile
    const int rstride = hcols * EKP_HEAT_CH;
    const int c = (int) fastdiv(task, magic_of(nstrips));
    const int strip = task - c * nstrips;
    nms_task_at<kDebug, EKP_HEAT_CH>(p, g.img, c, strip, m0, tb, X0, TW, sHeat + c - g.hr0 * rstride - g.hc0 * EKP_HEAT_CH, rstride, sTaps);
}

// The calling warp takes surviving NMS tasks from the tile's list until none is left.
template <bool kDebug>
__device__ __forceinline__ void nms_task_loop(const DenseParams& p, const TileGeom& g, const float* sHeat, const TileCtl& ctl) {
    const int nactive = *ctl.num_active;
    for (;;) {
        int item = 0;
        if ((threadIdx.x & 31) == 0) item = atomicAdd(ctl.next_task, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nactive) break;
        nms_task<kDebug>(p, g, sHeat, ctl.taps, ctl.list[item]);
    }
}

// ---- one tile without materialisation (the lean and the debug kernel) ----------------------------------
template <bool kDebug>
__device__ __forceinline__ void process_tile_lean(const DenseParams& p, const TileGeom& g, const TileSmem& sm, const TileCtl& ctl) {
    const int hcols = p.tile_wl + 6;
    stage_patch<EKP_HEAT_CH>(sm.heat, p.heat, p.layout, g.img, p.h, p.w, g.hr0, g.hr1, g.hc0, g.hc1, hcols, threadIdx.x, kLeanThreads);
    stage_commit();
    stage_wait<0>();
    __syncthreads();
    if (!kDebug) {
        build_colmax(g, sm.heat, hcols, ctl.colmax, threadIdx.x, kLeanThreads);
        __syncthreads();
    }
    build_task_list<kDebug>(p, g, sm.heat, ctl, threadIdx.x, kLeanThreads);
    __syncthreads();
    nms_task_loop<kDebug>(p, g, sm.heat, ctl);
}

// ---- one tile with materialisation: TMA bulk stores, two warp roles -----------------------------------
//   fill warps (kFillWarps): stage the PAF patch, turn it into paf_mat chunks in the store buffers and
//       hand them to the TMA engine; then (once the other group has signalled BAR_HEAT_READY) the
//       same for heat_mat; then help with the NMS tasks;
//   NMS warps (the rest): stage the heat patch, build the early-out table and the task list, signal
//       BAR_HEAT_READY, run the smoothing + NMS tasks.
// The two groups only meet at that one barrier, so the smoothing of a tile runs underneath its stores
// and the TMA queue of the SM (kStoreBufs chunks per resident CTA) never runs dry for long.
__device__ __forceinline__ void process_tile_mat(const DenseParams& p, const TileGeom& g, const TileSmem& sm, const TileCtl& ctl) {
    const int img = g.img, m0 = g.m0, i0 = g.i0, twl = g.twl, tb = g.tb;
    const int h = p.h, w = p.w, H = 8 * h, W = 8 * w;
    const int hcols = p.tile_wl + 6, pcols = p.tile_wl + 2;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp < kFillWarps) {
        const int t = threadIdx.x;
        stage_patch<EKP_PAF_CH>(sm.paf, p.paf, p.layout, img, h, w, g.pr0, g.pr1, g.pc0, g.pc1, pcols, t, kFillThreads);
        stage_commit();
        stage_wait<0>();
        bar_sync(BAR_FILL, kFillThreads);
        int phase = 0;
        materialise_tile<EKP_PAF_CH>(sm.paf, g.pr0, g.pc0, pcols, p.paf_mat + (size_t) img * H * W * EKP_PAF_CH, h, w, m0, tb, i0,
                                          twl, sm.store, phase, t);
        bar_sync(BAR_HEAT_READY, kMatThreads);  // heat patch staged, task list built
        if (p.heat_mat)
            materialise_tile<EKP_HEAT_CH>(sm.heat, g.hr0, g.hc0, hcols, p.heat_mat + (size_t) img * H * W * EKP_HEAT_CH, h, w, m0,
                                               tb, i0, twl, sm.store, phase, t);
        nms_task_loop<false>(p, g, sm.heat, ctl);
        // shared memory must outlive the TMA engine's reads of the store buffers
        if ((t & 31) == 0) bulk_wait_read<0>();
    } else {
        const int t = threadIdx.x - kFillThreads;
        stage_patch<EKP_HEAT_CH>(sm.heat, p.heat, p.layout, img, h, w, g.hr0, g.hr1, g.hc0, g.hc1, hcols, t, kNmsThreads);
        stage_commit();
        stage_wait<0>();
        bar_sync(BAR_NMS, kNmsThreads);
        build_colmax(g, sm.heat, hcols, ctl.colmax, t, kNmsThreads);
        bar_sync(BAR_NMS, kNmsThreads);
        build_task_list<false>(p, g, sm.heat, ctl, t, kNmsThreads);
        __threadfence_block();
        bar_sync(BAR_NMS, kNmsThreads);     // list and count complete for this group ...
        bar_arrive(BAR_HEAT_READY, kMatThreads);  // ... and published to the fill warps
        nms_task_loop<false>(p, g, sm.heat, ctl);
    }
}

template <bool kMat, bool kDebug, int TB>
__global__ void __launch_bounds__(kMat ? kMatThreads : kLeanThreads, kMat ? kMatBlocks : kLeanBlocks)
dense_frontend_kernel(const DenseParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(16) float sTaps[64];
    __shared__ unsigned short sList[kMaxSubTiles * EKP_NUM_PART * ((8 * kMaxTwl + 29) / 30 + 1)];
    __shared__ int sNumActive, sNextTask;
    if (threadIdx.x < 64) sTaps[threadIdx.x] = cTapsInterior[threadIdx.x >> 3][threadIdx.x & 7];
    if (threadIdx.x == 0) { sNumActive = 0; sNextTask = 0; }
    const int hcols = p.tile_wl + 6, pcols = p.tile_wl + 2;
    TileSmem sm;
    sm.heat = smem;
    sm.paf = sm.heat + (TB + 6) * hcols * EKP_HEAT_CH;
    TileCtl ctl;
    ctl.colmax = sm.paf + (kMat ? kPafRows * pcols * EKP_PAF_CH : 0);  // no PAF patch without materialisation
    ctl.taps = sTaps; ctl.list = sList; ctl.num_active = &sNumActive; ctl.next_task = &sNextTask;
    // store buffers behind the patches, 128-byte aligned (the patches' size is a multiple of 4 bytes only)
    sm.store = reinterpret_cast<float4*>(smem + (((size_t) (ctl.colmax + hcols * EKP_HEAT_CH - smem) + 31) & ~(size_t) 31));
    const TileGeom g = tile_geom<TB>(p, blockIdx.x, blockIdx.y, blockIdx.z);
    if (!kDebug) {
        // The first wave of CTAs pulls the WHOLE batch's stride-8 inputs into L2 (evict_last) in one burst
        // before the write stream builds up: reads that trickle in between 2.3 GB of stores cost far more than
        // their 36 MB (HBM read/write turnarounds).  One thread per CTA hands its contiguous slice of each
        // tensor to the TMA engine (cp.async.bulk.prefetch.L2); measured against per-line prefetch instructions
        // by all threads (-1 %) and no prefetch (-2.3 %), profiles/README.md.
        const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        const unsigned ncta = gridDim.x * gridDim.y * gridDim.z;
        constexpr unsigned kPrefetchCtas = 148u * (kMat ? kMatBlocks : kLeanBlocks);  // one resident wave on a B200
        const unsigned nfirst = ncta < kPrefetchCtas ? ncta : kPrefetchCtas;
        if (lin < nfirst && threadIdx.x == 0) {
            unsigned long long policy;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
            auto prefetch_slice = [&](const float* base, size_t bytes) {
                // 16-byte aligned pieces of [base, base + bytes): a hint, so the few bytes around a misaligned end do not matter
                const uintptr_t lo = (reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t) 15;
                const uintptr_t hi = (reinterpret_cast<uintptr_t>(base) + bytes) & ~(uintptr_t) 15;
                if (hi <= lo) return;
                const size_t per = (((size_t) (hi - lo) + nfirst - 1) / nfirst + 15) & ~(size_t) 15;
                const uintptr_t a0 = lo + (size_t) lin * per;
                if (a0 >= hi) return;
                const unsigned sz = (unsigned) (a0 + per < hi ? per : hi - a0);
                asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(a0), "r"(sz), "l"(policy));
            };
            prefetch_slice(p.heat, (size_t) p.n * EKP_HEAT_CH * p.h * p.w * 4);
            if (kMat) prefetch_slice(p.paf, (size_t) p.n * EKP_PAF_CH * p.h * p.w * 4);
        }
    }
    __syncthreads();  // sTaps and the task counters are initialised
    if (kMat) process_tile_mat(p, g, sm, ctl);
    else process_tile_lean<kDebug>(p, g, sm, ctl);
}

// ---- the plane kernel: stages 1-3 without materialisation for the network's own layout (NCHW) ----------------------
// One CTA per (part, image).  In NCHW a part's stride-8 map is ONE contiguous plane of h*w floats (10-60 KB at the
// BASELINE shapes): the TMA engine brings it into shared memory with a single bulk copy (cp.async.bulk global ->
// shared, completion on an mbarrier) -- no halo re-staging (the tiled kernel stages every sample ~4.9 times), no index
// arithmetic, no transposition.  Then, per CTA:
//   * a table M[row][strip] = max(0, maximum of the stride-8 row over the <= 9 columns a 30-column strip can touch);
//   * the exact early-out of build_task_list per (row-block pair, strip) task from 5 table entries per row block;
//   * the surviving tasks run nms_task_at on the plane (column stride 1), pulled from a list by the CTA's warps.
// Same arithmetic, same peaks as the tiled kernels (tests/test_gpu_parity.py); 43 M -> ~21 M warp instructions on the
// 64 x 368x432 batch (the tiled kernel spent ~60 % of its instructions on staging and the early-out tables).
constexpr int kPlaneThreads = 256;

// A plane may be cut into `slices` groups of row-block pairs (blockIdx.z), each staging only its rows + halo -- still
// one contiguous range of the plane -- so that a small batch of big maps fills the GPU (16 x 1312x736: 288 -> 864 CTAs).
__host__ __device__ inline int plane_pairs(int h) { return (h + kTaskTB - 1) / kTaskTB; }
__host__ __device__ inline void plane_slice_rows(int h, int slices, int slice, int& rp_lo, int& rp_hi, int& r_lo, int& r_hi) {
    const int nrp = plane_pairs(h);
    if (slices == 1) { rp_lo = 0; rp_hi = nrp; }   // (the usual case: no 64-bit divisions in every thread)
    else {
        rp_lo = (int) ((long long) nrp * slice / slices);
        rp_hi = (int) ((long long) nrp * (slice + 1) / slices);
    }
    const int m0 = rp_lo * kTaskTB, m1 = min(rp_hi * kTaskTB, h);   // row blocks [m0, m1)
    r_lo = max(min(m0 - 3, h - 5), 0);                             // staged rows: the windows of the rows and their halo rows
    r_hi = min(max(m1 + 2, 4), h - 1);
}
static size_t plane_smem_bytes(int h, int w, int slices) {
    const int nstrips = (8 * w + 29) / 30;
    int rows = 0, pairs = 0;
    for (int sl = 0; sl < slices; sl++) {
        int a, b, r0, r1;
        plane_slice_rows(h, slices, sl, a, b, r0, r1);
        rows = r1 - r0 + 1 > rows ? r1 - r0 + 1 : rows;
        pairs = b - a > pairs ? b - a : pairs;
    }
    size_t bytes = sizeof(float) * (size_t) ((rows * w + 3) & ~3);     // the staged rows of the plane
    bytes += sizeof(float) * (size_t) rows * nstrips;                  // M[row][strip]
    bytes += sizeof(unsigned short) * (size_t) ((pairs * nstrips + 1) & ~1);  // surviving tasks
    return bytes;
}

__global__ void __launch_bounds__(kPlaneThreads) dense_plane_kernel(const DenseParams p, int slices) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(16) float sTaps[64];
    __shared__ __align__(8) unsigned long long sBar;
    __shared__ int sNumActive, sNextTask;
    const int c = blockIdx.x, img = blockIdx.y;
    const int h = p.h, w = p.w, W = 8 * w;
    const int nstrips = (W + 29) / 30;
    int rp_lo, rp_hi, r_lo, r_hi;
    plane_slice_rows(h, slices, blockIdx.z, rp_lo, rp_hi, r_lo, r_hi);
    const int nrows = r_hi - r_lo + 1, nel = nrows * w, ntask = (rp_hi - rp_lo) * nstrips;
    float* sRows = smem;
    float* sM = sRows + ((nel + 3) & ~3);
    unsigned short* sList = reinterpret_cast<unsigned short*>(sM + nrows * nstrips);
    const float* src = p.heat + ((size_t) img * EKP_HEAT_CH + c) * h * w + (size_t) r_lo * w;
    const bool bulk = (nel & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;  // 16-byte aligned source and size
    if (threadIdx.x == 0) {
        sNumActive = 0; sNextTask = 0;
        if (bulk) mbar_init(&sBar, 1);
    }
    if (threadIdx.x < 64) sTaps[threadIdx.x] = cTapsInterior[threadIdx.x >> 3][threadIdx.x & 7];
    if (bulk) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&sBar, (unsigned) nel * 4u);
            bulk_load(sRows, src, (unsigned) nel * 4u, &sBar);
        }
        mbar_wait(&sBar, 0);
    } else {
        for (int i = threadIdx.x; i < nel; i += kPlaneThreads) sRows[i] = __ldg(src + i);
        __syncthreads();
    }
    const float* base = sRows - r_lo * w;   // base[j * w + i] = sample (row j, column i) of the map, j in [r_lo, r_hi]
    // idx / nstrips as one multiply-high: exact while idx * nstrips < 2^32 (both are below 2^16: the task list holds shorts)
    const unsigned div_ns = 0xffffffffu / (unsigned) nstrips + 1u;
    // M[r][s]: what row r can contribute to strip s at most (the columns a strip's lanes read: build_task_list)
    for (int idx = threadIdx.x; idx < nrows * nstrips; idx += kPlaneThreads) {
        const int r = (int) __umulhi((unsigned) idx, div_ns), strip = idx - r * nstrips;
        const int Xa = -1 + 30 * strip;
        const int xlo = min(max(Xa, 0), W - 1), xhi = min(max(Xa + 31, 0), W - 1);
        const int c_lo = min(max((xlo >> 3) - 2, 0), w - 5), c_hi = min(max((xhi >> 3) - 2, 0), w - 5) + 4;
        float mx = 0.f;
        const float* rowp = sRows + r * w;
#pragma unroll
        for (int k = 0; k < 9; k++) mx = fmaxf(mx, rowp[min(c_lo + k, c_hi)]);   // 5..9 columns (a strip spans <= 5 stride-8 cells + 4): straight-line code
        sM[idx] = mx;
    }
    __syncthreads();
    // exact early-out per (row-block pair, strip): same bounds as build_task_list
    for (int t = threadIdx.x; t < ntask; t += kPlaneThreads) {
        const int q = (int) __umulhi((unsigned) t, div_ns);
        const int rp = rp_lo + q, strip = t - q * nstrips;
        const int b_lo = rp * kTaskTB, b_hi = min(b_lo + kTaskTB, h);
        bool active = !(p.thr > 0.f);
        for (int m = b_lo; m < b_hi && !active; m++) {
            const int wb = min(max(m - 2, 0), h - 5);
            const bool interior = m >= 2 && m <= h - 3;
            float bound = 0.f;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const float mj = sM[(wb + j - r_lo) * nstrips + strip];
                bound = interior ? fmaf(cTapsInteriorMax[j], mj, bound) : fmaxf(bound, mj);
            }
            active = bound > p.thr * 0.9999f;
        }
        if (active) sList[atomicAdd(&sNumActive, 1)] = (unsigned short) t;
    }
    __syncthreads();
    const int nactive = sNumActive;
    for (;;) {
        int item = 0;
        if ((threadIdx.x & 31) == 0) item = atomicAdd(&sNextTask, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nactive) break;
        const int t = sList[item];
        const int q = (int) __umulhi((unsigned) t, div_ns);
        const int rp = rp_lo + q, strip = t - q * nstrips;
        const int m0 = rp * kTaskTB;
        nms_task_at<false, 1>(p, img, c, strip, m0, min(kTaskTB, h - m0), 0, W, base, w, sTaps);
    }
}

static size_t smem_bytes(int tile_wl, bool materialise, int tb) {
    size_t floats = (size_t) (tb + 6) * (tile_wl + 6) * EKP_HEAT_CH + (size_t) (tile_wl + 6) * EKP_HEAT_CH;
    if (!materialise) return sizeof(float) * floats;
    floats += (size_t) kPafRows * (tile_wl + 2) * EKP_PAF_CH;
    floats = (floats + 31) & ~(size_t) 31;
    return sizeof(float) * floats + sizeof(float4) * (size_t) kStoreBufs * kStoreBufF4;
}
size_t dense_frontend_smem_bytes(int tile_wl, bool materialise) { return smem_bytes(tile_wl, materialise, materialise ? kTB : kLeanTB); }
// choose the stride-8 tile width: <= kMaxTwl columns, tiles of (nearly) equal width
int dense_frontend_tile_wl(int w) {
    const int nt = (w + kMaxTwl - 1) / kMaxTwl;
    return (w + nt - 1) / nt;
}

// per device, once (ekp_create): allow the largest tile's dynamic shared memory
cudaError_t configure_dense_frontend() {
    cudaError_t e = raise_dynamic_smem_limit(dense_frontend_kernel<true, false, kTB>, smem_bytes(kMaxTwl, true, kTB));
    if (e == cudaSuccess) e = raise_dynamic_smem_limit(dense_frontend_kernel<false, false, kTB>, smem_bytes(kMaxTwl, false, kTB));
    if (e == cudaSuccess) e = raise_dynamic_smem_limit(dense_frontend_kernel<false, false, kLeanTB>, smem_bytes(kMaxTwl, false, kLeanTB));
    if (e == cudaSuccess) e = raise_dynamic_smem_limit(dense_frontend_kernel<false, true, kTB>, smem_bytes(kMaxTwl, false, kTB));
    return e;
}
// per context: the plane kernel's shared memory depends on the largest map the context takes
constexpr size_t kPlaneSmemMax = 200 * 1024;
cudaError_t configure_dense_plane(int max_h, int max_w) {
    const size_t need = plane_smem_bytes(max_h, max_w, 1);
    return raise_dynamic_smem_limit(dense_plane_kernel, need <= kPlaneSmemMax ? need : 48 * 1024);
}

// One CTA per tile.  (Persistent variants -- resident CTAs pulling tiles from a counter with the next tile's
// patches prefetched into a second buffer -- were measured SLOWER on B200 twice: 0.442 vs 0.412 ms with
// per-thread stores, 0.408 vs 0.381 ms with the bulk stores and warp roles; see profiles/README.md.)
static bool lean_tiled_forced() {  // EKP_LEAN_TILED=1: the tiled lean kernel also for NCHW input (measurements)
    static const bool v = getenv("EKP_LEAN_TILED") && atoi(getenv("EKP_LEAN_TILED")) != 0;
    return v;
}
cudaError_t launch_dense_frontend(const DenseParams& p, cudaStream_t stream) {
    const unsigned gx = (p.w + p.tile_wl - 1) / p.tile_wl;
    auto grid = [&](int tb) { return dim3(gx, (p.h + tb - 1) / tb, p.n); };
    if (p.smooth_out) {
        dense_frontend_kernel<false, true, kTB><<<grid(kTB), kLeanThreads, smem_bytes(p.tile_wl, false, kTB), stream>>>(p);
    } else if (p.paf_mat) {
        dense_frontend_kernel<true, false, kTB><<<grid(kTB), kMatThreads, smem_bytes(p.tile_wl, true, kTB), stream>>>(p);
    } else if (p.layout == EKP_LAYOUT_NCHW && plane_smem_bytes(p.h, p.w, 1) <= kPlaneSmemMax && ((8 * p.w + 29) / 30) * plane_pairs(p.h) < 65536 &&
               !lean_tiled_forced()) {
        // Lean, NCHW (the network's layout): CTAs per (part, image[, slice of rows]) on the part's contiguous plane;
        // a batch too small to give every SM a CTA (a single frame: 18 planes) is sliced until there are about two per SM
        // (slicing a launch that already fills the GPU only adds halo rows: 16 x 1312x736 99 -> 107 us, measured)
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int slices = EKP_NUM_PART * p.n >= sms ? 1 : (2 * sms + EKP_NUM_PART * p.n - 1) / (EKP_NUM_PART * p.n);
        slices = slices < 1 ? 1 : (slices > plane_pairs(p.h) / 2 ? (plane_pairs(p.h) / 2 > 0 ? plane_pairs(p.h) / 2 : 1) : slices);
        if (slices > 16) slices = 16;
        dense_plane_kernel<<<dim3(EKP_NUM_PART, p.n, slices), kPlaneThreads, plane_smem_bytes(p.h, p.w, slices), stream>>>(p, slices);
    } else {
        // Lean, other layouts / huge maps: 16-row tiles balance best while there are few of them; with many waves of tiles taller ones win (their
        // fixed cost -- staging a 6-row halo, the early-out tables -- is shared by twice the rows).  Same results either way.
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const dim3 g2 = grid(kTB);
        const bool tall = kLeanTB != kTB && (size_t) g2.x * g2.y * g2.z >= (size_t) 8 * sms * kLeanBlocks;
        if (tall) dense_frontend_kernel<false, false, kLeanTB><<<grid(kLeanTB), kLeanThreads, smem_bytes(p.tile_wl, false, kLeanTB), stream>>>(p);
        else dense_frontend_kernel<false, false, kTB><<<g2, kLeanThreads, smem_bytes(p.tile_wl, false, kTB), stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace ekp
