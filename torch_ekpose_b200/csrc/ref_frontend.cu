// ref_frontend.cu -- the REFERENCE's own peak front-end as CUDA kernels (sm_100a).
//
// Replaces, for a whole batch, the Python side of the reference's post-processing
// (/root/reference/lib/utils/paf_to_pose.py):
//   find_peaks :26-36   stride-8 NMS with the 4-neighbour cross footprint (scipy maximum_filter,
//                       mode='reflect' == in-bounds neighbours only), value > THRESH_HEATMAP;
//   NMS :60-133         per peak: clip a (<=5)x(<=5) window, cv2.resize(fx=8, fy=8, INTER_CUBIC),
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

__global__ void __launch_bounds__(kRefWarps * 32) ref_frontend_kernel(const RefParams p) {
    __shared__ float sPatch[kRefWarps][25];
    __shared__ float sTmp[kRefWarps][5 * 40];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = p.h, w = p.w;
    const long long task = (long long) blockIdx.x * kRefWarps + warp;  // (img, part, row)
    if (task >= (long long) p.n * EKP_NUM_PART * h) return;
    const int y = (int) (task % h);
    const int part = (int) ((task / h) % EKP_NUM_PART);
    const int img = (int) (task / ((long long) h * EKP_NUM_PART));

    for (int xb = 0; xb < w; xb += 32) {
        const int x = xb + lane;
        bool is_max = false;
        if (x < w) {
            const float v = lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x);
            is_max = v > p.thr;
            if (is_max && x > 0) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x - 1) > v);
            if (is_max && x < w - 1) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x + 1) > v);
            if (is_max && y > 0) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y - 1, x) > v);
            if (is_max && y < h - 1) is_max = !(lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y + 1, x) > v);
        }
        unsigned mask = __ballot_sync(0xffffffffu, is_max);
        if (!p.refine) {  // bool_refine_center=False: the maxima themselves, at (c + 0.5) * 8 - 0.5 truncated, heat value as score
            if (is_max) {
                const int slot = atomicAdd(p.raw_count + img, 1);
                if (slot < p.raw_cap) {
                    RawPeak pk;
                    pk.x = 8 * x + 3;
                    pk.y = 8 * y + 3;
                    pk.score = lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y, x);
                    pk.part = part;
                    pk.key = ((unsigned) y << 16) | (unsigned) x;
                    p.raw[(size_t) img * p.raw_cap + slot] = pk;
                }
            }
            continue;
        }
        while (mask) {  // the whole warp refines one peak at a time
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int px = xb + src;
            const int x_min = max(px - 2, 0), x_max = min(px + 2, w - 1);
            const int y_min = max(y - 2, 0), y_max = min(y + 2, h - 1);
            const int ph = y_max - y_min + 1, pw = x_max - x_min + 1;
            const int W8 = pw * 8, H8 = ph * 8;
            __syncwarp();
            if (lane < ph * pw) {
                const int r = lane / pw, c = lane - r * pw;
                sPatch[warp][r * 5 + c] = lo_at(p.heat, p.layout, img, EKP_HEAT_CH, h, w, part, y_min + r, x_min + c);
            }
            __syncwarp();
            // horizontal pass (HResizeCubic): tmp[r][dx]
            for (int idx = lane; idx < ph * W8; idx += 32) {
                const int r = idx / W8, dx = idx - r * W8;
                const int q = dx + 4;
                const int sx = (q >> 3) - 1;
                const float* a = p.cubic + (q & 7) * 4;
                const float* S = sPatch[warp] + r * 5;
                float v = __fmul_rn(S[min(max(sx - 1, 0), pw - 1)], __ldg(a + 0));
                v = __fadd_rn(v, __fmul_rn(S[min(max(sx, 0), pw - 1)], __ldg(a + 1)));
                v = __fadd_rn(v, __fmul_rn(S[min(max(sx + 1, 0), pw - 1)], __ldg(a + 2)));
                v = __fadd_rn(v, __fmul_rn(S[min(max(sx + 2, 0), pw - 1)], __ldg(a + 3)));
                sTmp[warp][r * 40 + dx] = v;
            }
            __syncwarp();
            // vertical pass (VResizeCubicVec_32f order) fused with the arg-max
            float best = -INFINITY;
            int best_idx = 0x7fffffff;
            for (int idx = lane; idx < H8 * W8; idx += 32) {
                const int dy = idx / W8, dx = idx - dy * W8;
                const int q = dy + 4;
                const int sy = (q >> 3) - 1;
                const float* b = p.cubic + (q & 7) * 4;
                const float* T = sTmp[warp] + dx;
                float v = __fmul_rn(T[min(max(sy + 2, 0), ph - 1) * 40], __ldg(b + 3));
                v = __fadd_rn(__fmul_rn(T[min(max(sy + 1, 0), ph - 1) * 40], __ldg(b + 2)), v);
                v = __fadd_rn(__fmul_rn(T[min(max(sy, 0), ph - 1) * 40], __ldg(b + 1)), v);
                v = __fadd_rn(__fmul_rn(T[min(max(sy - 1, 0), ph - 1) * 40], __ldg(b + 0)), v);
                if (v > best || best_idx == 0x7fffffff) { best = v; best_idx = idx; }  // first maximum
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
                const int slot = atomicAdd(p.raw_count + img, 1);
                if (slot < p.raw_cap) {
                    RawPeak pk;
                    pk.x = 8 * x_min + (best_idx % W8);
                    pk.y = 8 * y_min + (best_idx / W8);
                    pk.score = best;
                    pk.part = part;
                    pk.key = ((unsigned) y << 16) | (unsigned) px;
                    p.raw[(size_t) img * p.raw_cap + slot] = pk;
                }
            }
        }
    }
}

cudaError_t launch_ref_frontend(const RefParams& p, cudaStream_t stream) {
    const long long tasks = (long long) p.n * EKP_NUM_PART * p.h;
    const unsigned grid = (unsigned) ((tasks + kRefWarps - 1) / kRefWarps);
    ref_frontend_kernel<<<grid, kRefWarps * 32, 0, stream>>>(p);
    return cudaGetLastError();
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
