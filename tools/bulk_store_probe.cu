// Probe: can the materialise stores go faster through the TMA engine (cp.async.bulk shared -> global)
// than through per-warp st.global.cs.v4?  Same tile decomposition as dense_frontend_kernel
// (TWL stride-8 columns x 16 rows per CTA, 64 images of 368x432, heat_mat 19 ch + paf_mat 38 ch),
// constant data, no other work.
//   mode 0: per-thread 16-byte streaming stores, column outer / rows inner (the kernel's order)
//   mode 1: every output row segment is first written to shared memory by all threads
//           (st.shared.v4), then ONE thread hands it to the TMA engine as one bulk copy
//           (16 KB heat / 32 KB paf per row at TWL=27); NBUF row buffers in flight
//   mode 2: as 1 but without refilling shared memory (pure bulk-store floor)
//   mode 3: warp-private: a warp owns a 32-float4 (512-byte) column chunk like in the kernel, fills
//           8 rows x 512 B of its own shared-memory buffer, then lane 0 issues 8 bulk copies of 512 B
//           (no block barrier; two buffers per warp)
//   mode 4: CTA-wide chunks in the kernel's own thread mapping: thread = float4 column, 8 rows inner;
//           a buffer holds [8 rows][THREADS columns]; thread 0 issues 8 bulk copies of THREADS x 16 B
//   mode 5: as 4, but threads 0, 32, 64, .. (one per warp) issue one row each
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_store_probe tools/bulk_store_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int H = 368, W = 432, N = 64, TB = 2;

// HINT=1 (nvcc -DHINT=1): L2 evict_first policy on the bulk stores
#ifndef HINT
#define HINT 0
#endif
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
#if HINT
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"((unsigned) __cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned) __cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory"); }
template <int K>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(K) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int C, int TWL, int MODE, int NBUF, int THREADS>
__device__ __forceinline__ void tile_store(float* out_img, int m0, int i0, float val, float4* sbuf, int& phase) {
    const int X0 = i0 * 8;
    constexpr int row_f4 = TWL * 2 * C;
    const size_t stride4 = (size_t) W * C / 4;
    const float4 v = make_float4(val, val + 1, val + 2, val + 3);
    if (MODE == 0) {
        for (int col = threadIdx.x; col < row_f4; col += THREADS) {
            float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) (8 * m0) * W + X0) * C) + col;
#pragma unroll
            for (int k = 0; k < 8 * TB; k++) { __stcs(dst, v); dst += stride4; }
        }
    } else if (MODE == 4 || MODE == 5) {
        for (int c0 = 0; c0 < row_f4; c0 += THREADS) {
            const int ncol = min(THREADS, row_f4 - c0);
            for (int half = 0; half < TB; half++) {
                float4* buf = sbuf + (size_t) (phase % NBUF) * (8 * THREADS);
                if (MODE == 4 ? threadIdx.x == 0 : (threadIdx.x & 31) == 0) bulk_wait_read<NBUF - 1>();
                __syncthreads();
                if ((int) threadIdx.x < ncol) {
#pragma unroll
                    for (int k = 0; k < 8; k++) buf[k * THREADS + threadIdx.x] = v;
                }
                fence_async();
                __syncthreads();
                float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) (8 * (m0 + half)) * W + X0) * C) + c0;
                if (MODE == 4) {
                    if (threadIdx.x == 0) {
#pragma unroll
                        for (int k = 0; k < 8; k++) bulk_store(dst + (size_t) k * stride4, buf + k * THREADS, ncol * 16);
                        bulk_commit();
                    }
                } else if ((threadIdx.x & 31) == 0) {
                    for (int k = threadIdx.x >> 5; k < 8; k += THREADS / 32) bulk_store(dst + (size_t) k * stride4, buf + k * THREADS, ncol * 16);
                    bulk_commit();
                }
                phase++;
            }
        }
    } else if (MODE == 3) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        float4* wbuf = sbuf + (size_t) warp * (2 * 8 * 32);
        for (int chunk = warp; chunk * 32 < row_f4; chunk += THREADS / 32) {
            const int ncol = min(32, row_f4 - chunk * 32);
            for (int half = 0; half < TB; half++) {
                float4* buf = wbuf + (size_t) (phase & 1) * (8 * 32);
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; k++) buf[k * 32 + lane] = v;
                fence_async();
                __syncwarp();
                if (lane == 0) {
                    float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) (8 * (m0 + half)) * W + X0) * C) + chunk * 32;
#pragma unroll
                    for (int k = 0; k < 8; k++) bulk_store(dst + (size_t) k * stride4, buf + k * 32, ncol * 16);
                    bulk_commit();
                }
                phase++;
            }
        }
    } else {
        constexpr int buf_f4 = TWL * 2 * 38;  // buffers sized for the PAF row
        for (int k = 0; k < 8 * TB; k++) {
            float4* buf = sbuf + (size_t) (phase % NBUF) * buf_f4;
            // the bulk copy that last read this buffer must have finished READING shared memory
            if (threadIdx.x == 0) bulk_wait_read<NBUF - 1>();
            __syncthreads();
            if (MODE == 1)
                for (int col = threadIdx.x; col < row_f4; col += THREADS) buf[col] = v;
            fence_async();
            __syncthreads();
            if (threadIdx.x == 0) {
                float4* dst = reinterpret_cast<float4*>(out_img + ((size_t) (8 * m0 + k) * W + X0) * C);
                bulk_store(dst, buf, row_f4 * 16);
                bulk_commit();
            }
            phase++;
        }
    }
}

template <int TWL, int MODE, int NBUF, int THREADS>
__global__ void __launch_bounds__(THREADS) pattern_kernel(float* heat_mat, float* paf_mat) {
    extern __shared__ __align__(128) float4 dyn[];
    const int img = blockIdx.z, m0 = blockIdx.y * TB, i0 = blockIdx.x * TWL;
    int phase = 0;
    tile_store<38, TWL, MODE, NBUF, THREADS>(paf_mat + (size_t) img * H * W * 38, m0, i0, (float) img, dyn, phase);
    tile_store<19, TWL, MODE, NBUF, THREADS>(heat_mat + (size_t) img * H * W * 19, m0, i0, (float) img, dyn, phase);
    if (MODE != 0 && (threadIdx.x & 31) == 0) bulk_wait<0>();
}

__global__ void linear_kernel(float4* out, size_t n4) {
    const float4 v = make_float4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) __stcs(out + i, v);
}

static float *hm, *pm;
static cudaEvent_t ea, eb;
static double total_bytes;

template <int TWL, int MODE, int NBUF, int THREADS>
void run(const char* label, int ctas_per_sm) {
    auto k = pattern_kernel<TWL, MODE, NBUF, THREADS>;
    const size_t need = MODE == 0 ? 0 : MODE == 3 ? (size_t) (THREADS / 32) * 2 * 8 * 32 * 16 : MODE >= 4 ? (size_t) NBUF * 8 * THREADS * 16 : (size_t) NBUF * TWL * 2 * 38 * 16;
    size_t smem = need;
    if (ctas_per_sm > 0) {  // limit residency through dynamic shared memory
        const size_t cap = (size_t) (224 * 1024 / ctas_per_sm) - 2048;
        if (cap < need) { printf("%-58s skipped (needs %zu B)\n", label, need); return; }
        smem = cap;
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    dim3 grid(54 / TWL, 23, N);
    float best = 1e9, sum = 0;
    int cnt = 0;
    for (int it = 0; it < 14; it++) {
        cudaEventRecord(ea);
        k<<<grid, THREADS, smem>>>(hm, pm);
        cudaEventRecord(eb); cudaEventSynchronize(eb);
        float ms; cudaEventElapsedTime(&ms, ea, eb);
        if (it >= 4) { best = ms < best ? ms : best; sum += ms; cnt++; }
    }
    cudaError_t e = cudaGetLastError();
    printf("%-58s ctas/sm<=%d  best %.3f ms  mean %.3f ms  %.0f GB/s  %s\n", label, ctas_per_sm, best, sum / cnt, total_bytes / best / 1e6,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const size_t nh = (size_t) N * H * W * 19, np = (size_t) N * H * W * 38;
    cudaMalloc(&hm, nh * 4); cudaMalloc(&pm, np * 4);
    cudaEventCreate(&ea); cudaEventCreate(&eb);
    total_bytes = (double) (nh + np) * 4;
    {
        float best = 1e9;
        for (int it = 0; it < 12; it++) {
            cudaEventRecord(ea);
            linear_kernel<<<148 * 8, 256>>>((float4*) pm, np / 4); linear_kernel<<<148 * 8, 256>>>((float4*) hm, nh / 4);
            cudaEventRecord(eb); cudaEventSynchronize(eb);
            float ms; cudaEventElapsedTime(&ms, ea, eb);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("%-58s best %.3f ms  %.0f GB/s\n", "linear st.global.cs.v4 (2 launches)", best, total_bytes / best / 1e6);
        best = 1e9;
        for (int it = 0; it < 12; it++) {
            cudaEventRecord(ea);
            cudaMemsetAsync(pm, 0, np * 4); cudaMemsetAsync(hm, 0, nh * 4);
            cudaEventRecord(eb); cudaEventSynchronize(eb);
            float ms; cudaEventElapsedTime(&ms, ea, eb);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("%-58s best %.3f ms  %.0f GB/s\n", "cudaMemsetAsync (2 calls)", best, total_bytes / best / 1e6);
    }
    for (int occ : {2, 3, 4}) run<27, 0, 1, 256>("st.global.cs.v4, TWL 27, 256 thr", occ);
    for (int occ : {1, 2, 3}) run<54, 0, 1, 256>("st.global.cs.v4, TWL 54, 256 thr", occ);
    for (int occ : {1, 2, 3}) run<27, 1, 2, 256>("bulk (fill smem + TMA), TWL 27, 2 bufs, 256 thr", occ);
    for (int occ : {1, 2}) run<27, 1, 3, 256>("bulk (fill smem + TMA), TWL 27, 3 bufs, 256 thr", occ);
    for (int occ : {1, 2}) run<27, 1, 2, 512>("bulk (fill smem + TMA), TWL 27, 2 bufs, 512 thr", occ);
    for (int occ : {1, 2, 3}) run<27, 2, 2, 256>("bulk (TMA only), TWL 27, 2 bufs, 256 thr", occ);
    for (int occ : {1, 2, 3}) run<27, 3, 2, 256>("bulk warp-private 8 x 512 B, TWL 27, 256 thr", occ);
    for (int occ : {1, 2, 3}) run<27, 4, 2, 256>("bulk CTA chunk 8 x 4 KB, 1 issuer, 2 bufs, 256 thr", occ);
    for (int occ : {1, 2, 3}) run<27, 5, 2, 256>("bulk CTA chunk 8 x 4 KB, 8 issuers, 2 bufs, 256 thr", occ);
    for (int occ : {1, 2}) run<27, 4, 3, 256>("bulk CTA chunk 8 x 4 KB, 1 issuer, 3 bufs, 256 thr", occ);
    for (int occ : {1, 2, 3}) run<27, 4, 2, 192>("bulk CTA chunk 8 x 3 KB, 1 issuer, 2 bufs, 192 thr", occ);
    for (int occ : {1, 2}) run<27, 4, 2, 512>("bulk CTA chunk 8 x 8 KB, 1 issuer, 2 bufs, 512 thr", occ);
    for (int occ : {1, 2}) run<27, 5, 2, 512>("bulk CTA chunk 8 x 8 KB, 8 issuers, 2 bufs, 512 thr", occ);
    for (int occ : {1}) run<54, 1, 2, 512>("bulk (fill smem + TMA), TWL 54, 2 bufs, 512 thr", occ);
    for (int occ : {1}) run<54, 2, 2, 256>("bulk (TMA only), TWL 54, 2 bufs, 256 thr", occ);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
