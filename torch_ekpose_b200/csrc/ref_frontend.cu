// ref_frontend.cu -- the REFERENCE's own peak front-end as CUDA kernels (sm_100a).
//
// Replaces, for a whole batch, the Python side of the reference's post-processing
// (/root/reference/lib/utils/paf_to_pose.py):
//   find_peaks :26-36   stride-8 NMS with the 4-neighbour cross footprint (scipy maximum_filter,
//                       mode='reflect' == in-bounds neighbours only), value > THRESH_HEATMAP;
//   NMS :60-133         per peak: clip a (<=5)x(<=5) window, cv2.resize(fx=8, fy=8, INTER_CUBIC),
//                       [bool_gaussian_filt: scipy.ndimage.gaussian_filter(sigma=3) of that patch, :111-112,]
//                       first arg-max -> refined integer full-resolution coordinate and score;
//   paf_to_pose_cpp :356-359  cv2.resize(INTER_NEAREST) x8 of PAF / heat (upsample_nearest_kernel,
//                       only when the caller asks for the operator-surface tensors).
// The bicubic arithmetic is OpenCV's own float path (resize.cpp: interpolateCubic with A=-0.75,
// horizontal pass summed left to right, vertical pass b3*S3 + b2*S2 + b1*S1 + b0*S0 accumulated
// from S3, replicate border, no FMA) as restated in oracle/frontend_oracle.c, which is pinned
// bit-for-bit to cv2 4.13 with IPP off.  Peaks come out unordered; peaks_sort_kernel restores
// the reference's (part, y, x) order from RawPeak::key.
#include "common.cuh"

namespace ekp {

constexpr int kRefWarps = 8;

// cv2's bicubic coefficients for the eight x8 phases t = (2k + 1) / 16 (capi.cu build_cubic_table): the vertical pass of
// ref_refine_kernel reads them as constant-bank operands; the horizontal pass (a different phase per lane) keeps the copy
// in global memory (RefParams::cubic).
__constant__ float cCubic[8][4];
cudaError_t set_cubic_table(const float* tab32) { return cudaMemcpyToSymbol(cCubic, tab32, sizeof(float) * 32); }

// ---- find_peaks: one block per (part, image) ----------------------------------------------------------------------------
// Writes one RawPeak per maximum: with p.refine == 0 already in its final form (NMS(bool_refine_center=False): the
// maximum itself at (c + 0.5) * 8 - 0.5 truncated, heat value as score), else the stride-8 cell (x, y) for
// ref_refine_kernel to replace by the refined peak.
// NCHW (the network's layout): the part's map is one contiguous plane; the TMA engine brings it into shared memory with a
// single bulk copy (as in dense_plane_kernel) and the threads walk it linearly -- a third of the instructions of
// per-element global indexing.  Other layouts / planes that do not fit read global memory directly.
__device__ __forceinline__ void ref_emit(const RefParams& p, int img, int part, int x, int y, float v, bool is_max) {
    const unsigned mask = __ballot_sync(0xffffffffu, is_max);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(mask) - 1) base = atomicAdd(p.raw_count + img, __popc(mask));   // one atomic per warp and step
    base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
    if (is_max) {
        const int slot = base + __popc(mask & ((1u << lane) - 1u));
        if (slot < p.raw_cap) {
            RawPeak pk;
            pk.x = p.refine ? x : 8 * x + 3;
            pk.y = p.refine ? y : 8 * y + 3;
            pk.score = v;
            pk.part = part;
            pk.key = ((unsigned) y << 16) | (unsigned) x;
            p.raw[(size_t) img * p.raw_cap + slot] = pk;
        }
    }
}

template <bool kPlane>
__global__ void __launch_bounds__(kRefWarps * 32) ref_scan_kernel(const RefParams p) {
    extern __shared__ __align__(16) float sPlane[];
    __shared__ __align__(8) unsigned long long sBar;
    const int h = p.h, w = p.w, hw = h * w;
    const int part = blockIdx.x, img = blockIdx.y;
    if (kPlane) {
        const float* src = p.heat + ((size_t) img * EKP_HEAT_CH + part) * hw;
        const bool bulk = (hw & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        if (bulk) {
            if (threadIdx.x == 0) mbar_init(&sBar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_expect_tx(&sBar, (unsigned) hw * 4u);
                bulk_load(sPlane, src, (unsigned) hw * 4u, &sBar);
            }
            mbar_wait(&sBar, 0);
        } else {
            for (int i = threadIdx.x; i < hw; i += kRefWarps * 32) sPlane[i] = __ldg(src + i);
            __syncthreads();
        }
        const int steps = (hw + kRefWarps * 32 - 1) / (kRefWarps * 32);
        int y = threadIdx.x / w, x = threadIdx.x - y * w;   // (row, column) of element threadIdx.x, advanced without divisions
        const int dy = (kRefWarps * 32) / w, dx = (kRefWarps * 32) - dy * w;
        for (int s = 0, i = threadIdx.x; s < steps; s++, i += kRefWarps * 32) {
            bool is_max = false;
            float v = 0.f;
            if (i < hw) {
                v = sPlane[i];
                is_max = v > p.thr && !(x > 0 && sPlane[i - 1] > v) && !(x < w - 1 && sPlane[i + 1] > v) &&
                         !(y > 0 && sPlane[i - w] > v) && !(y < h - 1 && sPlane[i + w] > v);
            }
            ref_emit(p, img, part, x, y, v, is_max);
            x += dx; y += dy;
            if (x >= w) { x -= w; y++; }
        }
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int y = warp; y < h; y += kRefWarps)
            for (int xb = 0; xb < w; xb += 32) {
                const int x = xb + lane;
                bool is_max = false;
                float v = 0.f;
                if (x < w) {
                    v = lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x);
                    is_max = v > p.thr;
                    if (is_max && x > 0) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x - 1) > v);
                    if (is_max && x < w - 1) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x + 1) > v);
                    if (is_max && y > 0) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y - 1, x) > v);
                    if (is_max && y < h - 1) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y + 1, x) > v);
                }
                ref_emit(p, img, part, x, y, v, is_max);
            }
    }
}

// ---- NMS(): bicubic refinement, one warp per peak, all peaks of the batch at once ------------------------------------
// (Refining inside the scan -- the warp that found a row's maxima refined them one after the other -- left most of the
// GPU waiting for the few warps whose rows held peaks: 64 x 368x432 62 -> 57 us even with the cheaper inner loop below.)
// kGauss: NMS(bool_gaussian_filt=True).  The upsampled patch goes through scipy.ndimage.gaussian_filter(sigma=3) before
// the arg-max: radius 12, one correlate1d per axis (axis 0 first, float32 in between), 'reflect' extension, and because
// the kernel is symmetric NI_Correlate1D evaluates  tmp = x[i] w[0];  tmp += (x[i+jj] + x[i-jj]) w[jj], jj = -12 .. -1
// in double -- the same operations here (oracle/frontend_oracle.c okp_scipy_gauss3, pinned bit-for-bit to scipy).
constexpr int kGaussPatch = 40 * 40;   // floats per buffer; two buffers per warp (dynamic shared memory)
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - 1 - i : i); }  // radius 12 < n

template <bool kGauss>
__global__ void __launch_bounds__(kRefWarps * 32) ref_refine_kernel(const RefParams p) {
    extern __shared__ __align__(16) float sGaussBuf[];   // kGauss: [kRefWarps][2][kGaussPatch]
    __shared__ float sPatch[kRefWarps][25];
    __shared__ float sTmp[kRefWarps][5 * 40];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sV = sGaussBuf + (size_t) warp * 2 * kGaussPatch;   // the upsampled patch, row stride 40
    float* sT = sV + kGaussPatch;                              // ... after the axis-0 pass
    const int h = p.h, w = p.w, img = blockIdx.y;
    const int count = min(p.raw_count[img], p.raw_cap);
    for (int slot = blockIdx.x * kRefWarps + warp; slot < count; slot += gridDim.x * kRefWarps) {   // (uniform per warp)
    __syncwarp();
    RawPeak* out = p.raw + (size_t) img * p.raw_cap + slot;
    const int px = out->x, y = out->y, part = out->part;
    const int x_min = max(px - 2, 0), x_max = min(px + 2, w - 1);
    const int y_min = max(y - 2, 0), y_max = min(y + 2, h - 1);
    const int ph = y_max - y_min + 1, pw = x_max - x_min + 1;
    const int W8 = pw * 8, H8 = ph * 8;
    if (lane < ph * pw) {
        const int r = lane / pw, c = lane - r * pw;
        sPatch[warp][r * 5 + c] = lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y_min + r, x_min + c);
    }
    __syncwarp();
    // Lane = output column (two slots: dx = lane and dx = 32 + lane, a patch is at most 40 columns wide), so the
    // column's source indices and coefficients are per-lane constants and the row's are warp-uniform: no index
    // arithmetic per output value.
    // horizontal pass (HResizeCubic): tmp[r][dx]
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
        const int dx = 32 * sl + lane;
        if (dx < W8) {
            const int q = dx + 4;
            const int sx = (q >> 3) - 1;
            const float* a = p.cubic + (q & 7) * 4;
            const float a0 = __ldg(a + 0), a1 = __ldg(a + 1), a2 = __ldg(a + 2), a3 = __ldg(a + 3);
            const int c0 = min(max(sx - 1, 0), pw - 1), c1 = min(max(sx, 0), pw - 1), c2 = min(max(sx + 1, 0), pw - 1),
                      c3 = min(max(sx + 2, 0), pw - 1);
            for (int r = 0; r < ph; r++) {
                const float* S = sPatch[warp] + r * 5;
                float v = __fmul_rn(S[c0], a0);
                v = __fadd_rn(v, __fmul_rn(S[c1], a1));
                v = __fadd_rn(v, __fmul_rn(S[c2], a2));
                v = __fadd_rn(v, __fmul_rn(S[c3], a3));
                sTmp[warp][r * 40 + dx] = v;
            }
        }
    }
    __syncwarp();
    // vertical pass (VResizeCubicVec_32f order) fused with the arg-max: first maximum in row-major order.  Eight output rows
    // per source row: output row 8 yb + k has phase (k + 4) & 7 and reads the four rows around sy = yb - 1 (k < 4) or yb
    // (k >= 4), so inside the unrolled block the coefficients are constant-bank operands and the row pointers are set up
    // twice per eight rows.  A lane meets its elements in increasing row-major index (row by row, slot 0 before slot 1), so
    // "first maximum" within a lane is a strict > against what it holds; the lanes' results are merged with the index
    // tie-break below.  (Before: coefficients loaded per row, pointers per row, a three-way test per element: 12.2 M -> see
    // profiles/README.md.)
    float best = -INFINITY;
    int best_idx = lane < W8 ? lane : 0x7fffffff;   // the first element (row 0) stands until something is greater
    for (int yb = 0; yb < ph; yb++) {
        const float* Tm2 = sTmp[warp] + min(max(yb - 2, 0), ph - 1) * 40;
        const float* Tm1 = sTmp[warp] + min(max(yb - 1, 0), ph - 1) * 40;
        const float* Tz = sTmp[warp] + yb * 40;
        const float* Tp1 = sTmp[warp] + min(yb + 1, ph - 1) * 40;
        const float* Tp2 = sTmp[warp] + min(yb + 2, ph - 1) * 40;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int phs = (k + 4) & 7;
            // sy = yb - 1 (k < 4): rows sy - 1 .. sy + 2 = yb - 2 .. yb + 1;  sy = yb (k >= 4): yb - 1 .. yb + 2
            const float* T0 = k < 4 ? Tm2 : Tm1;
            const float* T1 = k < 4 ? Tm1 : Tz;
            const float* T2 = k < 4 ? Tz : Tp1;
            const float* T3 = k < 4 ? Tp1 : Tp2;
            const int dy = 8 * yb + k;
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int dx = 32 * sl + lane;
                if (dx < W8) {
                    float v = __fmul_rn(T3[dx], cCubic[phs][3]);
                    v = __fadd_rn(__fmul_rn(T2[dx], cCubic[phs][2]), v);
                    v = __fadd_rn(__fmul_rn(T1[dx], cCubic[phs][1]), v);
                    v = __fadd_rn(__fmul_rn(T0[dx], cCubic[phs][0]), v);
                    if (kGauss) { sV[dy * 40 + dx] = v; continue; }
                    if (v > best) { best = v; best_idx = dy * W8 + dx; }
                }
            }
        }
    }
    if (kGauss) {
        double fw[13];
#pragma unroll
        for (int j = 0; j < 13; j++) fw[j] = __ldg(p.gauss + j);   // fw[12 + jj] = w[jj]
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {   // axis 0 (along y): lane = column
            const int dx = 32 * sl + lane;
            if (dx < W8) {
                for (int dy = 0; dy < H8; dy++) {
                    double tmp = __dmul_rn((double) sV[dy * 40 + dx], fw[12]);
#pragma unroll
                    for (int jj = -12; jj < 0; jj++) {
                        const double a = (double) sV[reflect1(dy + jj, H8) * 40 + dx], b = (double) sV[reflect1(dy - jj, H8) * 40 + dx];
                        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), fw[12 + jj]));
                    }
                    sT[dy * 40 + dx] = __double2float_rn(tmp);
                }
            }
        }
        __syncwarp();
        for (int dy = 0; dy < H8; dy++) {   // axis 1 (along x) fused with the arg-max
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const int dx = 32 * sl + lane;
                if (dx < W8) {
                    const float* T = sT + dy * 40;
                    double tmp = __dmul_rn((double) T[dx], fw[12]);
#pragma unroll
                    for (int jj = -12; jj < 0; jj++) {
                        const double a = (double) T[reflect1(dx + jj, W8)], b = (double) T[reflect1(dx - jj, W8)];
                        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), fw[12 + jj]));
                    }
                    const float v = __double2float_rn(tmp);
                    const int idx = dy * W8 + dx;
                    if (v > best || (v == best && idx < best_idx) || best_idx == 0x7fffffff) { best = v; best_idx = idx; }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
        if (oi != 0x7fffffff && (best_idx == 0x7fffffff || ov > best || (ov == best && oi < best_idx))) {
            best = ov; best_idx = oi;
        }
    }
    if (lane == 0) {
        out->x = 8 * x_min + (best_idx % W8);
        out->y = 8 * y_min + (best_idx / W8);
        out->score = best;
    }
    }
}

constexpr size_t kGaussBytes = sizeof(float) * 2 * kGaussPatch * kRefWarps;
cudaError_t launch_ref_frontend(const RefParams& p, cudaStream_t stream) {
    const size_t plane_bytes = sizeof(float) * (size_t) ((p.h * p.w + 3) & ~3);
    if (p.layout == EKP_LAYOUT_NCHW && plane_bytes <= 200 * 1024 && p.w <= kRefWarps * 32)
        ref_scan_kernel<true><<<dim3(EKP_NUM_PART, p.n), kRefWarps * 32, plane_bytes, stream>>>(p);
    else
        ref_scan_kernel<false><<<dim3(EKP_NUM_PART, p.n), kRefWarps * 32, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !p.refine) return e;
    // the peak counts are only known on the device: a fixed number of blocks per image (about four resident blocks per SM
    // over the batch), whose warps stride over the image's peak list
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_image = (4 * sms + p.n - 1) / p.n;
    per_image = per_image < 1 ? 1 : (per_image > (p.raw_cap + kRefWarps - 1) / kRefWarps ? (p.raw_cap + kRefWarps - 1) / kRefWarps : per_image);
    if (p.refine == 2) ref_refine_kernel<true><<<dim3(per_image, p.n), kRefWarps * 32, kGaussBytes, stream>>>(p);
    else ref_refine_kernel<false><<<dim3(per_image, p.n), kRefWarps * 32, 0, stream>>>(p);
    return cudaGetLastError();
}
int ref_frontend_launches(int refine) { return refine ? 2 : 1; }
cudaError_t configure_ref_frontend(int max_h, int max_w) {
    const size_t plane_bytes = sizeof(float) * (size_t) ((max_h * max_w + 3) & ~3);
    cudaError_t e = raise_dynamic_smem_limit(ref_scan_kernel<true>, plane_bytes <= 200 * 1024 ? plane_bytes : 48 * 1024);
    if (e == cudaSuccess) e = raise_dynamic_smem_limit(ref_refine_kernel<true>, kGaussBytes);
    return e;
}

// ---- nearest x8 upsample (paf_to_pose.py:356-359), HWC output --------------------------------
// One block per (image, stride-8 row): stage the row as HWC in shared memory, expand it x8 along x into a
// shared-memory row buffer (<= kUpChunk stride-8 columns at a time), and let the TMA engine write it to the 8
// identical full-resolution rows (cp.async.bulk shared -> global, one copy of up to 58 KB per output row,
// L2 evict_first like the dense front-end's stores).
constexpr int kUpChunk = 48;
template <int C>
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const float* __restrict__ lo, int layout, int h, int w,
                                                               float* __restrict__ out) {
    extern __shared__ __align__(128) float sUp[];  // [8 * cw * C] expanded chunk, then [w][C] staged row
    const int j = blockIdx.x, img = blockIdx.y;
    const int cw_max = w < kUpChunk ? w : kUpChunk;
    float* sOut = sUp;
    float* sRow = sUp + (size_t) 8 * cw_max * C;
    if (layout == EKP_LAYOUT_NCHW) {
        for (int idx = threadIdx.x; idx < C * w; idx += 256) {
            const int c = idx / w, i = idx - c * w;
            sRow[i * C + c] = __ldg(lo + (((size_t) img * C + c) * h + j) * w + i);
        }
    } else {
        for (int idx = threadIdx.x; idx < C * w; idx += 256) sRow[idx] = __ldg(lo + (((size_t) img * h + j) * w) * C + idx);
    }
    __syncthreads();
    const int W = 8 * w;
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    for (int i0 = 0; i0 < w; i0 += cw_max) {
        const int cw = min(cw_max, w - i0);
        const int n4 = 8 * cw * C / 4;
        for (int col = threadIdx.x; col < n4; col += 256) {
            float e[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int f = col * 4 + k;
                const int x = f / C, c = f - x * C;
                e[k] = sRow[(i0 + (x >> 3)) * C + c];
            }
            reinterpret_cast<float4*>(sOut)[col] = make_float4(e[0], e[1], e[2], e[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 8) {  // one output row each
            float* dst = out + (((size_t) img * 8 * h + 8 * j + threadIdx.x) * W + 8 * i0) * C;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
                         "r"((unsigned) __cvta_generic_to_shared(sOut)), "r"((unsigned) (n4 * 16)), "l"(policy)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the buffer is refilled (or the block exits) next
        }
        __syncthreads();
    }
}

static size_t upsample_smem(int w, int C) {
    const int cw = w < kUpChunk ? w : kUpChunk;
    return sizeof(float) * ((size_t) 8 * cw * C + (size_t) w * C);
}
cudaError_t launch_upsample_nearest(const float* lo, int layout, int n, int h, int w, int C, float* out, cudaStream_t stream) {
    dim3 grid(h, n);
    const size_t smem = upsample_smem(w, C);
    cudaError_t e;
    if (C == EKP_PAF_CH) {
        e = raise_dynamic_smem_limit(upsample_nearest_kernel<EKP_PAF_CH>, smem);
        if (e != cudaSuccess) return e;
        upsample_nearest_kernel<EKP_PAF_CH><<<grid, 256, smem, stream>>>(lo, layout, h, w, out);
    } else if (C == EKP_HEAT_CH) {
        e = raise_dynamic_smem_limit(upsample_nearest_kernel<EKP_HEAT_CH>, smem);
        if (e != cudaSuccess) return e;
        upsample_nearest_kernel<EKP_HEAT_CH><<<grid, 256, smem, stream>>>(lo, layout, h, w, out);
    } else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace ekp
