// Probe: how fast can B200 absorb the materialise store pattern with no compute at all?
// Same grid / tile / thread->address mapping as dense_frontend_kernel's materialise_tile
// (2 x 23 x 64 CTAs of 256 threads, each thread a float4 column, 16 rows inner), constant data.
// Variants: rows-inner (the kernel's order), 2 columns interleaved, linear (reference).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int H = 368, W = 432, N = 64, TWL = 27, TB = 2;

template <int C, int MODE>
__device__ __forceinline__ void tile_store(float* out_img, int m0, int i0, float val) {
    const int X0 = i0 * 8;
    const int row_f4 = TWL * 2 * C;
    const size_t stride4 = (size_t) W * C / 4;
    const float4 v = make_float4(val, val + 1, val + 2, val + 3);
    if (MODE == 0) {  // kernel order: column outer, 16 rows inner
        for (int col = threadIdx.x; col < row_f4; col += 256) {
            float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) (8 * m0) * W + X0) * C) + col;
#pragma unroll
            for (int k = 0; k < 8 * TB; k++) { __stcs(dst, v); dst += stride4; }
        }
    } else {  // row outer: the CTA sweeps each row segment completely before the next row
        for (int k = 0; k < 8 * TB; k++) {
            float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) (8 * m0 + k) * W + X0) * C);
            for (int col = threadIdx.x; col < row_f4; col += 256) __stcs(dst + col, v);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) pattern_kernel(float* heat_mat, float* paf_mat) {
    extern __shared__ float dyn[];
    if (threadIdx.x == 999) dyn[0] = 0.f;
    const int img = blockIdx.z, m0 = blockIdx.y * TB, i0 = blockIdx.x * TWL;
    tile_store<38, MODE>(paf_mat + (size_t) img * H * W * 38, m0, i0, (float) img);
    tile_store<19, MODE>(heat_mat + (size_t) img * H * W * 19, m0, i0, (float) img);
}

__global__ void linear_kernel(float4* out, size_t n4) {
    const float4 v = make_float4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) __stcs(out + i, v);
}

int main() {
    const size_t nh = (size_t) N * H * W * 19, np = (size_t) N * H * W * 38;
    float *hm, *pm;
    cudaMalloc(&hm, nh * 4); cudaMalloc(&pm, np * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    dim3 grid(2, 23, N);
    const double bytes = (double) (nh + np) * 4;
    cudaFuncSetAttribute(pattern_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int occ = 8; occ >= 1; occ--) {  // limit resident CTAs per SM through dynamic shared memory
        const size_t smem = occ == 8 ? 0 : (size_t) (220 * 1024 / occ) - 1024;
        float best = 1e9;
        for (int it = 0; it < 10; it++) {
            cudaEventRecord(a);
            pattern_kernel<0><<<grid, 256, smem>>>(hm, pm);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("pattern, <= %d CTAs/SM (%zu B smem): %.3f ms  %.1f GB/s\n", occ, smem, best, bytes / best / 1e6);
    }
    for (int variant = 0; variant < 3; variant++) {
        float best = 1e9;
        for (int it = 0; it < 12; it++) {
            cudaEventRecord(a);
            if (variant == 0) pattern_kernel<0><<<grid, 256>>>(hm, pm);
            else if (variant == 1) pattern_kernel<1><<<grid, 256>>>(hm, pm);
            else { linear_kernel<<<148 * 8, 256>>>((float4*) pm, np / 4); linear_kernel<<<148 * 8, 256>>>((float4*) hm, nh / 4); }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("variant %d (%s): %.3f ms  %.1f GB/s\n", variant, variant == 0 ? "col-outer rows-inner" : variant == 1 ? "row-outer" : "linear", best, bytes / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
