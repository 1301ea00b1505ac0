// dense_frontend.cu -- stages 1-3 of the hot path, fused, for sm_100a.
//
// One launch takes the network's stride-8 heat (19 ch) and PAF (38 ch) maps of a whole batch
// and, per 16-row x (8*tile_wl)-column full-resolution tile,
//   (1) stages the stride-8 neighbourhood of the tile in shared memory (HWC order),
//   (2) optionally materialises the bilinear x8 tensors heat_mat[H][W][19] / paf_mat[H][W][38]
//       (the operator-surface tensors of process_paf, paf_to_pose.py:356-360) with 16-byte
//       coalesced streaming stores -- this is the HBM-bound part: 36.25 MB written per 368x432
//       image against 0.57 MB read,
//   (3) evaluates the Gaussian-smoothed (sigma 3, 25 taps, reflect) bilinear-upsampled heat map
//       as ONE separable 5-tap polyphase filter on the stride-8 grid (the composition of the two
//       linear operators; tables from host, see dense_tables.cpp) entirely in registers,
//   (4) does the 3x3 max NMS with warp shuffles (x) and a rolling 3-row window (y) and appends
//       peaks with a warp-ballot aggregated atomic into the per-image raw peak list.
// Nothing but the (optional) operator-surface tensors and the peaks ever goes back to HBM.
//
// Arithmetic is the one defined in oracle/frontend_oracle.c part (B); results are bit-identical
// to it (tests/test_gpu_parity.py).  There is no reference implementation of this front-end
// (SURVEY.md 0.1); the reference's own front-end is ref_frontend.cu.
#include "common.cuh"

namespace ekp {

constexpr int kThreads = 256;
constexpr int kTH = 16;             // full-resolution rows per tile
constexpr int kTB = kTH / 8;        // stride-8 row blocks per tile
constexpr int kHeatRows = kTB + 6;  // stride-8 rows staged for the smoothing window (+-3)
constexpr int kPafRows = kTB + 2;   // rows staged for bilinear only (+-1)

// ---- materialise one tensor (C channels) of the tile -----------------------------------------
// sP: HWC patch in shared memory with origin (pr0, pc0) and `pcols` columns per row.
// Every thread owns float4 columns of the tile's output rows: 4 consecutive (x, c) entries.
template <int C>
__device__ __forceinline__ void materialise_tile(const float* __restrict__ sP, int pr0, int pc0, int pcols,
                                                 float* __restrict__ out_img, int h, int w, int m0, int tb,
                                                 int i0, int twl) {
    const int W = w * 8;
    const int X0 = i0 * 8;
    const int row_f4 = twl * 2 * C;  // (8*twl*C)/4 float4 per tile row
    for (int col = threadIdx.x; col < row_f4; col += kThreads) {
        int off0[4], off1[4];
        float tx[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int f = col * 4 + e;
            const int xl = f / C;
            const int c = f - xl * C;
            int a, b;
            bilin_coord(X0 + xl, w, a, b, tx[e]);
            off0[e] = (a - pc0) * C + c;
            off1[e] = (b - pc0) * C + c;
        }
        float top[4], bot[4];
        {   // row pair (m0-1, m0) feeds rows 8*m0 .. 8*m0+3
            const float* r = sP + (size_t) (max(m0 - 1, 0) - pr0) * pcols * C;
#pragma unroll
            for (int e = 0; e < 4; e++) bot[e] = lerp1(r[off0[e]], r[off1[e]], tx[e]);
        }
        float* dst = out_img + ((size_t) (8 * m0) * W + X0) * C + (size_t) col * 4;
        for (int q = 0; q <= tb; q++) {
            // pair (m0+q-1, m0+q): rows 8*(m0+q)-4 .. 8*(m0+q)+3, clipped to the tile
            const float* r = sP + (size_t) (min(m0 + q, h - 1) - pr0) * pcols * C;
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                top[e] = bot[e];
                bot[e] = lerp1(r[off0[e]], r[off1[e]], tx[e]);
                d[e] = __fsub_rn(bot[e], top[e]);
            }
            const int k_lo = (q == 0) ? 4 : 0;   // first pair: only its lower half is in the tile
            const int k_hi = (q == tb) ? 4 : 8;  // last pair: only its upper half
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k >= k_lo && k < k_hi) {
                    const float ty = (float) (2 * k + 1) * 0.0625f;
                    float4 v;
                    v.x = fmaf(ty, d[0], top[0]);
                    v.y = fmaf(ty, d[1], top[1]);
                    v.z = fmaf(ty, d[2], top[2]);
                    v.w = fmaf(ty, d[3], top[3]);
                    __stcs(reinterpret_cast<float4*>(dst), v);
                    dst += (size_t) W * C;
                }
            }
        }
    }
}

// ---- stage a stride-8 patch in shared memory as [row][col][C] --------------------------------
template <int C>
__device__ __forceinline__ void stage_patch(float* __restrict__ sP, const float* __restrict__ src, int layout,
                                            int img, int h, int w, int r0, int r1, int c0, int c1, int pcols) {
    const int nr = r1 - r0 + 1, nc = c1 - c0 + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (layout == EKP_LAYOUT_NCHW) {
        // one warp per (channel, row) segment: coalesced along w in global, stride C in shared
        for (int cr = warp; cr < C * nr; cr += kThreads / 32) {
            const int c = cr / nr;
            const int r = cr - c * nr;
            const float* g = src + (((size_t) img * C + c) * h + (r0 + r)) * w + c0;
            float* s = sP + (size_t) r * pcols * C + c;
            for (int i = lane; i < nc; i += 32) s[i * C] = __ldg(g + i);
        }
    } else {
        const int per_r = nc * C;  // contiguous in both spaces
        for (int r = 0; r < nr; r++) {
            const float* g = src + (((size_t) img * h + (r0 + r)) * w + c0) * C;
            float* s = sP + (size_t) r * pcols * C;
            for (int k = threadIdx.x; k < per_r; k += kThreads) s[k] = __ldg(g + k);
        }
    }
}

__global__ void __launch_bounds__(kThreads) dense_frontend_kernel(const DenseParams p) {
    extern __shared__ __align__(16) float smem[];
    const int img = blockIdx.z;
    const int m0 = blockIdx.y * kTB;
    const int i0 = blockIdx.x * p.tile_wl;
    const int h = p.h, w = p.w, H = 8 * h, W = 8 * w;
    const int twl = min(p.tile_wl, w - i0);
    const int tb = min(kTB, h - m0);
    const int hcols = p.tile_wl + 6, pcols = p.tile_wl + 2;

    float* sHeat = smem;                                          // [kHeatRows][hcols][19]
    float* sPaf = sHeat + kHeatRows * hcols * EKP_HEAT_CH;        // [kPafRows][pcols][38]
    float* sAy = sPaf + kPafRows * pcols * EKP_PAF_CH;            // [kTH + 2][8]

    // the smoothing window of row block m is rows clamp(m-2, 0, h-5) .. +4 (same for columns)
    const int hr0 = max(min(m0 - 3, h - 5), 0), hr1 = min(max(m0 + tb + 2, 4), h - 1);
    const int hc0 = max(min(i0 - 3, w - 5), 0), hc1 = min(max(i0 + twl + 2, 4), w - 1);
    const int pr0 = max(m0 - 1, 0), pr1 = min(m0 + tb, h - 1);
    const int pc0 = max(i0 - 1, 0), pc1 = min(i0 + twl, w - 1);
    const bool mat = p.paf_mat != nullptr;

    stage_patch<EKP_HEAT_CH>(sHeat, p.heat, p.layout, img, h, w, hr0, hr1, hc0, hc1, hcols);
    if (mat) stage_patch<EKP_PAF_CH>(sPaf, p.paf, p.layout, img, h, w, pr0, pr1, pc0, pc1, pcols);
    const int Ya = 8 * m0 - 1;  // first row the NMS pass evaluates (halo)
    for (int idx = threadIdx.x; idx < (kTH + 2) * 8; idx += kThreads) {
        const int Y = Ya + (idx >> 3);
        sAy[idx] = (Y >= 0 && Y < H) ? __ldg(p.ay + (size_t) Y * 8 + (idx & 7)) : 0.f;
    }
    __syncthreads();

    // ---- (2) operator-surface tensors ------------------------------------------------------
    if (mat) {
        materialise_tile<EKP_PAF_CH>(sPaf, pr0, pc0, pcols, p.paf_mat + (size_t) img * H * W * EKP_PAF_CH, h, w, m0, tb, i0, twl);
        if (p.heat_mat)
            materialise_tile<EKP_HEAT_CH>(sHeat, hr0, hc0, hcols, p.heat_mat + (size_t) img * H * W * EKP_HEAT_CH, h, w, m0, tb, i0, twl);
    }

    // ---- (3)+(4) smoothed map and NMS --------------------------------------------------------
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 8 * twl, X0 = 8 * i0;
    const int nstrips = (TW + 29) / 30;
    const float NEG_INF = __int_as_float(0xff800000);
    for (int task = warp; task < EKP_NUM_PART * nstrips; task += kThreads / 32) {
        const int c = task / nstrips;
        const int strip = task - c * nstrips;
        const int X = X0 - 1 + 30 * strip + lane;
        const bool inb = X >= 0 && X < W;
        const bool out_lane = lane >= 1 && lane <= 30 && X < X0 + TW && X < W;
        const int Xc = min(max(X, 0), min(W - 1, X0 + TW));  // lanes past the halo are never outputs
        const int bx = min(max((Xc >> 3) - 2, 0), w - 5);
        const float4 axv = __ldg(reinterpret_cast<const float4*>(p.ax + (size_t) Xc * 8));
        const float ax4 = __ldg(p.ax + (size_t) Xc * 8 + 4);
        const float* colp = sHeat + (size_t) (bx - hc0) * EKP_HEAT_CH + c;
        const int rstride = hcols * EKP_HEAT_CH;

        auto trow = [&](int j) -> float {  // horizontal 5-tap pass on stride-8 row j
            const float* s = colp + (size_t) (j - hr0) * rstride;
            float acc = __fmul_rn(axv.x, s[0]);
            acc = fmaf(axv.y, s[EKP_HEAT_CH], acc);
            acc = fmaf(axv.z, s[2 * EKP_HEAT_CH], acc);
            acc = fmaf(axv.w, s[3 * EKP_HEAT_CH], acc);
            acc = fmaf(ax4, s[4 * EKP_HEAT_CH], acc);
            return acc;
        };

        float T0 = 0.f, T1 = 0.f, T2 = 0.f, T3 = 0.f, T4 = 0.f;
        int cur_wb = -100;
        float hm_prev2 = NEG_INF, hm_prev = NEG_INF, s_prev = NEG_INF;
        const int Yend = 8 * (m0 + tb);  // one row past the tile (halo)
        for (int Y = Ya; Y <= Yend; Y++) {
            float S = NEG_INF;
            if (Y >= 0 && Y < H) {  // uniform across the warp
                const int wb = min(max((Y >> 3) - 2, 0), h - 5);
                if (wb != cur_wb) {
                    if (wb == cur_wb + 1) {
                        T0 = T1; T1 = T2; T2 = T3; T3 = T4; T4 = trow(wb + 4);
                    } else {
                        T0 = trow(wb); T1 = trow(wb + 1); T2 = trow(wb + 2); T3 = trow(wb + 3); T4 = trow(wb + 4);
                    }
                    cur_wb = wb;
                }
                const float* ayr = sAy + (Y - Ya) * 8;
                const float4 a = *reinterpret_cast<const float4*>(ayr);
                const float a4 = ayr[4];
                float acc = __fmul_rn(a.x, T0);
                acc = fmaf(a.y, T1, acc);
                acc = fmaf(a.z, T2, acc);
                acc = fmaf(a.w, T3, acc);
                acc = fmaf(a4, T4, acc);
                if (inb) S = acc;
                if (p.smooth_out && out_lane && Y >= 8 * m0 && Y < Yend)
                    p.smooth_out[(((size_t) img * H + Y) * W + X) * EKP_NUM_PART + c] = acc;
            }
            const float l = __shfl_up_sync(0xffffffffu, S, 1);
            const float r = __shfl_down_sync(0xffffffffu, S, 1);
            const float hm = fmaxf(S, fmaxf(l, r));
            // row Y-1 is complete now: peak iff above threshold and equal to its 3x3 maximum
            const bool row_ok = (Y - 1) >= 8 * m0;  // tile-owned rows only (Y-1 < Yend always)
            const bool is_peak = row_ok && out_lane && s_prev > p.thr && s_prev == fmaxf(hm_prev2, fmaxf(hm_prev, hm));
            const unsigned mask = __ballot_sync(0xffffffffu, is_peak);
            if (mask) {
                const int leader = __ffs(mask) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(p.raw_count + img, __popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (is_peak) {
                    const int slot = base + __popc(mask & ((1u << lane) - 1u));
                    if (slot < p.raw_cap) {
                        RawPeak pk;
                        pk.x = X; pk.y = Y - 1; pk.score = s_prev; pk.part = c;
                        pk.key = ((unsigned) (Y - 1) << 16) | (unsigned) X;
                        p.raw[(size_t) img * p.raw_cap + slot] = pk;
                    }
                }
            }
            hm_prev2 = hm_prev; hm_prev = hm; s_prev = S;
        }
    }
}

size_t dense_frontend_smem_bytes(int tile_wl) {
    return sizeof(float) * ((size_t) kHeatRows * (tile_wl + 6) * EKP_HEAT_CH + (size_t) kPafRows * (tile_wl + 2) * EKP_PAF_CH +
                            (size_t) (kTH + 2) * 8);
}

// choose the stride-8 tile width: <= 32 columns, tiles of (nearly) equal width
int dense_frontend_tile_wl(int w) {
    const int nt = (w + 31) / 32;
    return (w + nt - 1) / nt;
}

// per device, once (ekp_create): allow the largest tile's dynamic shared memory
cudaError_t configure_dense_frontend() {
    return cudaFuncSetAttribute(dense_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int) dense_frontend_smem_bytes(32));
}

cudaError_t launch_dense_frontend(const DenseParams& p, cudaStream_t stream) {
    const size_t smem = dense_frontend_smem_bytes(p.tile_wl);
    dim3 grid((p.w + p.tile_wl - 1) / p.tile_wl, (p.h + kTB - 1) / kTB, p.n);
    dense_frontend_kernel<<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace ekp
