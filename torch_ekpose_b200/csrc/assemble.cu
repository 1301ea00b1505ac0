// assemble.cu -- the stand-alone launch of the assembly (assemble.cuh): one block per image.  The production path runs
// the same device code at the end of paf_connect_kernel (the block that finishes an image's last limb assembles it);
// this kernel serves per-stage timing and EKP_FUSE_ASSEMBLE=0.
#include "assemble.cuh"

namespace ekp {

#ifdef EKP_ASM_PROFILE
__device__ unsigned long long g_asm_prof[8];
extern "C" int ekp_debug_asm_profile(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, g_asm_prof, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_asm_prof, z, sizeof(z)); }
    return 0;
}
#endif

constexpr int kAsmThreads = 128;
__global__ void __launch_bounds__(kAsmThreads) assemble_kernel(const AsmParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    assemble_image(P, blockIdx.x, smem_raw);
}

cudaError_t configure_assemble(int max_humans, int max_peaks, int max_part) {
    return raise_dynamic_smem_limit(assemble_kernel, assemble_smem_bytes(max_humans, max_peaks, max_part,
                                                                         assemble_conn_cap(max_humans, max_part)));
}

cudaError_t launch_assemble(AsmParams P, int n, cudaStream_t stream) {
    P.conn_cap = assemble_conn_cap(P.max_humans, P.max_part);
    assemble_kernel<<<n, kAsmThreads, assemble_smem_bytes(P.max_humans, P.max_peaks, P.max_part, P.conn_cap), stream>>>(P);
    return cudaGetLastError();
}

}  // namespace ekp
