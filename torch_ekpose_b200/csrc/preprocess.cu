// preprocess.cu -- the input side of the path: the reference's
// `padding` (/root/reference/lib/evaluate/estimator.py:52-68: cv2.resize with the default
// INTER_LINEAR on the uint8 BGR frame so that the long side becomes dest_size, zero-pad to a
// multiple of `factor`) fused with `vgg_preprocess` / `rtpose_preprocess`
// (lib/datasets/preprocessing.py:16-43) into one kernel that turns a batch of equally sized
// uint8 frames into the network's float32 NCHW input on the device.
//
// cv2's 8-bit bilinear resize is OpenCV's 11-bit fixed-point algorithm (resize.cpp: coefficients
// cvRound(w * 2048) built on the host in capi.cu; horizontal pass into 32-bit, vertical pass
// (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2); the normalisation is IEEE float32 in the
// reference's operation order.  Bit-identical to the reference Python (oracle/frontend_oracle.c (C),
// tests/test_gpu_parity.py).
#include "common.cuh"

namespace ekp {

struct PreprocessParams {
    const unsigned char* src;  // [n][sh][sw][3] BGR
    float* out;                // [n][3][ph][pw]
    const int* xofs;           // [rw]
    const short* ialpha;       // [rw][2]
    const int* yofs;           // [rh]
    const short* ibeta;        // [rh][2]
    int n, sh, sw, rh, rw, ph, pw, mode;
};

__global__ void __launch_bounds__(128) preprocess_kernel(const PreprocessParams p) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= p.pw) return;
    int px[3] = {0, 0, 0};  // the resized-and-padded uint8 pixel, BGR
    if (x < p.rw && y < p.rh) {
        const int sx0 = p.xofs[x], sx1 = min(sx0 + 1, p.sw - 1);
        const int yo = p.yofs[y];
        const int sy0 = min(max(yo, 0), p.sh - 1), sy1 = min(max(yo + 1, 0), p.sh - 1);
        const int a0 = p.ialpha[2 * x], a1 = p.ialpha[2 * x + 1], b0 = p.ibeta[2 * y], b1 = p.ibeta[2 * y + 1];
        const unsigned char* s = p.src + (size_t) img * p.sh * p.sw * 3;
        const unsigned char* r0p = s + (size_t) sy0 * p.sw * 3;
        const unsigned char* r1p = s + (size_t) sy1 * p.sw * 3;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int r0 = r0p[sx0 * 3 + c] * a0 + r0p[sx1 * 3 + c] * a1;
            const int r1 = r1p[sx0 * 3 + c] * a0 + r1p[sx1 * 3 + c] * a1;
            const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            px[c] = min(max(v, 0), 255);
        }
    }
    const size_t plane = (size_t) p.ph * p.pw;
    float* o = p.out + (size_t) img * 3 * plane + (size_t) y * p.pw + x;
    if (p.mode == 0) {  // vgg_preprocess: /255, BGR -> RGB, (v - mean) / std
        const float means[3] = {0.485f, 0.456f, 0.406f}, stds[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
        for (int c = 0; c < 3; c++)
            o[c * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float) px[2 - c], 255.f), means[c]), stds[c]);
    } else {            // rtpose_preprocess: v / 256 - 0.5, channel order kept
#pragma unroll
        for (int c = 0; c < 3; c++) o[c * plane] = __fsub_rn(__fdiv_rn((float) px[c], 256.f), 0.5f);
    }
}

cudaError_t launch_preprocess(const unsigned char* src, float* out, const int* xofs, const short* ialpha, const int* yofs,
                              const short* ibeta, int n, int sh, int sw, int rh, int rw, int ph, int pw, int mode,
                              cudaStream_t stream) {
    PreprocessParams p;
    p.src = src; p.out = out; p.xofs = xofs; p.ialpha = ialpha; p.yofs = yofs; p.ibeta = ibeta;
    p.n = n; p.sh = sh; p.sw = sw; p.rh = rh; p.rw = rw; p.ph = ph; p.pw = pw; p.mode = mode;
    dim3 grid((pw + 127) / 128, ph, n);
    preprocess_kernel<<<grid, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace ekp
