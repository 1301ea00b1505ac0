// peaks_sort.cu -- peak ingest: from an unordered per-image peak list to the reference's
// part-sorted peak table (`peak_infos_line`, pafprocess.cpp:24-43) and per-part offsets.
//
//   peaks_ingest_kernel : process_paf's own input format, float [p2][p3] rows
//                         (x, y, score, _, part) -> RawPeak with key = input index
//                         (pafprocess.cpp:26-36: x,y truncated to int, id = running input index).
//   peaks_sort_kernel   : rank every peak by (part, key) -- keys are unique within a part, so the
//                         rank is a permutation and the result is independent of the order in which
//                         the front-end's atomics appended the peaks -- and scatter it to its row.
//                         id = rank for the front-ends (their input order IS the sorted order,
//                         paf_to_pose.py:350-352) or = input index for process_paf input, which
//                         reproduces the reference's indexing of peak_infos_line by id
//                         (pafprocess.cpp:208-218) including for unsorted input.
#include "common.cuh"

namespace ekp {

__global__ void peaks_ingest_kernel(const float* __restrict__ peaks, const int* __restrict__ n_peaks, int n_fixed,
                                    int peaks_stride, int p3, int W, int H, RawPeak* __restrict__ raw,
                                    int* __restrict__ raw_count, int raw_cap, unsigned* __restrict__ overflow) {
    const int img = blockIdx.y;
    const int n = n_peaks ? n_peaks[img] : n_fixed;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) raw_count[img] = n;  // peaks_sort clamps to raw_cap and flags the overflow
    if (k >= n || k >= raw_cap) return;
    const float* row = peaks + ((size_t) img * peaks_stride + k) * p3;
    RawPeak pk;
    pk.x = (int) row[0];  // C truncation, pafprocess.cpp:30-31
    pk.y = (int) row[1];
    pk.score = row[2];
    pk.part = (int) row[4];
    pk.key = (unsigned) k;
    if (pk.part < 0 || pk.part >= EKP_NUM_PART || pk.x < 0 || pk.x >= W || pk.y < 0 || pk.y >= H || !(pk.score == pk.score)) {
        atomicOr(overflow + img, EKP_OVF_BADPEAK);  // the reference has undefined behaviour here; we refuse
        pk.part = EKP_NUM_PART;                      // sorted behind every real part, never used
        pk.x = pk.y = 0;
    }
    raw[(size_t) img * raw_cap + k] = pk;
}

// Blocks (slice, image): every block buckets all of the image's 32-bit keys by part in shared memory (dynamic:
// raw_cap keys; a histogram, its prefix = the per-part offsets, then a scatter), and ranks the 256 peaks of its slice
// WITHIN their part's bucket: rank = part offset + number of smaller keys of the same part (keys are unique within a
// part, so the rank is a permutation whatever order the atomics filled the buckets in).  The inner loop runs over one
// part's peaks (tens) instead of all of the image's (hundreds in a crowd); slice 0 also writes the per-part offsets.
// Crowded images spread over several SMs (one block per 256 peaks).
constexpr int kSortThreads = 256;
__global__ void __launch_bounds__(kSortThreads) peaks_sort_kernel(const RawPeak* __restrict__ raw, const int* __restrict__ raw_count,
                                                                  int raw_cap, int max_part, int id_from_key, ekp_peak* __restrict__ line,
                                                                  int* __restrict__ part_off /* [n][20] */,
                                                                  int* __restrict__ n_peaks, unsigned* __restrict__ overflow) {
    extern __shared__ unsigned sKey[];                       // [raw_cap] keys, bucketed by part
    __shared__ int sCount[EKP_NUM_PART + 1], sBase[EKP_NUM_PART + 2], sFill[EKP_NUM_PART + 1];
    const int img = blockIdx.y, slice = blockIdx.x;
    const int total = raw_count[img];
    const int n = min(total, raw_cap);
    if (slice * kSortThreads >= n && slice > 0) return;  // nothing in this slice (slice 0 always writes the offsets)
    const RawPeak* r = raw + (size_t) img * raw_cap;
    if (threadIdx.x <= EKP_NUM_PART) { sCount[threadIdx.x] = 0; sFill[threadIdx.x] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kSortThreads) atomicAdd(&sCount[min(max(r[i].part, 0), EKP_NUM_PART)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int p = 0; p <= EKP_NUM_PART; p++) { sBase[p] = off; off += sCount[p]; }   // bucket 18: invalid peaks, sorted behind
        sBase[EKP_NUM_PART + 1] = off;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kSortThreads) {
        const int p = min(max(r[i].part, 0), EKP_NUM_PART);
        sKey[sBase[p] + atomicAdd(&sFill[p], 1)] = r[i].key;
    }
    __syncthreads();
    const int i = slice * kSortThreads + threadIdx.x;
    if (i < n) {
        const RawPeak pk = r[i];
        const int p = min(max(pk.part, 0), EKP_NUM_PART);
        const int lo = sBase[p], hi = sBase[p + 1];
        int rank = lo;
        for (int j = lo; j < hi; j++) rank += sKey[j] < pk.key;
        ekp_peak out;
        out.x = pk.x; out.y = pk.y; out.score = pk.score;
        out.id = id_from_key ? (int) pk.key : rank;
        line[(size_t) img * raw_cap + rank] = out;
    }
    if (slice != 0 || threadIdx.x != 0) return;
    unsigned ovf = 0;
    if (total > raw_cap) ovf |= EKP_OVF_PEAKS;
    int* po = part_off + (size_t) img * 20;
    for (int p = 0; p < EKP_NUM_PART; p++) {
        po[p] = sBase[p];
        if (sCount[p] > max_part) ovf |= EKP_OVF_PART;
    }
    po[EKP_NUM_PART] = sBase[EKP_NUM_PART];   // == number of valid peaks (invalid ones sort behind)
    po[EKP_NUM_PART + 1] = n;
    n_peaks[img] = sBase[EKP_NUM_PART];
    if (ovf) atomicOr(overflow + img, ovf);
}

cudaError_t launch_peaks_ingest(const float* peaks, const int* n_peaks, int n_fixed, int peaks_stride, int p3, int n,
                                int W, int H, RawPeak* raw, int* raw_count, int raw_cap, unsigned* overflow,
                                cudaStream_t stream) {
    const int maxn = min(peaks_stride, raw_cap);
    dim3 grid((maxn + 127) / 128 > 0 ? (maxn + 127) / 128 : 1, n);
    peaks_ingest_kernel<<<grid, 128, 0, stream>>>(peaks, n_peaks, n_fixed, peaks_stride, p3, W, H, raw, raw_count, raw_cap, overflow);
    return cudaGetLastError();
}

cudaError_t configure_peaks_sort(int raw_cap) {
    return raise_dynamic_smem_limit(peaks_sort_kernel, sizeof(unsigned) * (size_t) raw_cap);
}

cudaError_t launch_peaks_sort(const RawPeak* raw, const int* raw_count, int raw_cap, int max_part, int id_from_key, int n, ekp_peak* line,
                              int* part_off, int* n_peaks, unsigned* overflow, cudaStream_t stream) {
    const size_t smem = sizeof(unsigned) * (size_t) raw_cap;
    dim3 grid((raw_cap + kSortThreads - 1) / kSortThreads, n);
    peaks_sort_kernel<<<grid, kSortThreads, smem, stream>>>(raw, raw_count, raw_cap, max_part, id_from_key, line, part_off, n_peaks, overflow);
    return cudaGetLastError();
}

}  // namespace ekp
