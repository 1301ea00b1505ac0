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

// ---- one image, process_paf's input format, ingest + sort in ONE block (the host-pointer operator surface) --------
// peaks_ingest_kernel + peaks_sort_kernel for a single list of up to kOneMaxPeaks peaks without the intermediate
// RawPeak list and without the memset of the counters: the block validates the rows, buckets the input indices by
// part in shared memory, and writes every peak straight to its row of the part-sorted table (rank = part offset +
// number of earlier peaks of the same part: the reference's bucket order, pafprocess.cpp:24-43).  It also (re)sets the
// image's overflow word, so the call needs no cudaMemsetAsync in front.
constexpr int kOneThreads = 256;
constexpr int kOneMaxPeaks = 4096;
__global__ void __launch_bounds__(kOneThreads) peaks_ingest_sort_one_kernel(const float* __restrict__ peaks, int npk, const int* __restrict__ npk_dev,
                                                                            int p3, int W, int H, int raw_cap, int max_part,
                                                                            ekp_peak* __restrict__ line, int* __restrict__ part_off,
                                                                            int* __restrict__ n_peaks, int* __restrict__ raw_count,
                                                                            unsigned* __restrict__ overflow) {
    if (npk_dev) npk = max(*npk_dev, 0);   // the count travels with the peaks (a replayed CUDA graph has constant arguments)
    __shared__ unsigned short sKey[kOneMaxPeaks];   // input indices, bucketed by part
    __shared__ unsigned char sPart[kOneMaxPeaks];
    __shared__ int sCount[EKP_NUM_PART + 1], sBase[EKP_NUM_PART + 2], sFill[EKP_NUM_PART + 1];
    __shared__ unsigned sOvf;
    const int n = min(npk, min(raw_cap, kOneMaxPeaks));
    if (threadIdx.x <= EKP_NUM_PART) { sCount[threadIdx.x] = 0; sFill[threadIdx.x] = 0; }
    if (threadIdx.x == 0) sOvf = npk > raw_cap ? EKP_OVF_PEAKS : 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += kOneThreads) {
        const float* row = peaks + (size_t) k * p3;
        const int x = (int) row[0], y = (int) row[1], part = (int) row[4];
        const float sc = row[2];
        int p = part;
        if (part < 0 || part >= EKP_NUM_PART || x < 0 || x >= W || y < 0 || y >= H || !(sc == sc)) {
            atomicOr(&sOvf, EKP_OVF_BADPEAK);  // the reference has undefined behaviour here; we refuse
            p = EKP_NUM_PART;                  // sorted behind every real part, never used
        }
        sPart[k] = (unsigned char) p;
        atomicAdd(&sCount[p], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int p = 0; p <= EKP_NUM_PART; p++) { sBase[p] = off; off += sCount[p]; }
        sBase[EKP_NUM_PART + 1] = off;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += kOneThreads) sKey[sBase[sPart[k]] + atomicAdd(&sFill[sPart[k]], 1)] = (unsigned short) k;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += kOneThreads) {
        const int p = sPart[k];
        int rank = sBase[p];
        for (int j = sBase[p]; j < sBase[p + 1]; j++) rank += sKey[j] < k;
        const float* row = peaks + (size_t) k * p3;
        ekp_peak out;
        if (p < EKP_NUM_PART) { out.x = (int) row[0]; out.y = (int) row[1]; }
        else { out.x = 0; out.y = 0; }
        out.score = row[2];
        out.id = k;   // ids follow input order (pafprocess.cpp:29)
        line[rank] = out;
    }
    if (threadIdx.x != 0) return;
    unsigned ovf = sOvf;
    for (int p = 0; p < EKP_NUM_PART; p++) {
        part_off[p] = sBase[p];
        if (sCount[p] > max_part) ovf |= EKP_OVF_PART;
    }
    part_off[EKP_NUM_PART] = sBase[EKP_NUM_PART];
    part_off[EKP_NUM_PART + 1] = n;
    n_peaks[0] = sBase[EKP_NUM_PART];
    raw_count[0] = npk;
    overflow[0] = ovf;
}
int peaks_one_max() { return kOneMaxPeaks; }
cudaError_t launch_peaks_ingest_sort_one(const float* peaks, int npk, const int* npk_dev, int p3, int W, int H, int raw_cap, int max_part,
                                         ekp_peak* line, int* part_off, int* n_peaks, int* raw_count, unsigned* overflow, cudaStream_t stream) {
    peaks_ingest_sort_one_kernel<<<1, kOneThreads, 0, stream>>>(peaks, npk, npk_dev, p3, W, H, raw_cap, max_part, line, part_off, n_peaks,
                                                               raw_count, overflow);
    return cudaGetLastError();
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
