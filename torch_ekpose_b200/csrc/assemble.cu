// assemble.cu -- second half of stage 5: person assembly, pruning and the result tables, one
// warp per image.  Replaces /root/reference/lib/pafprocess/pafprocess.cpp:127-191 (subset
// assembly and pruning) and the getter loop of paf_to_pose_cpp (paf_to_pose.py:361-377).
//
// The assembly is inherently sequential over (limb, connection) and is kept so; only the row
// SEARCH (pafprocess.cpp:137-144) and the 18-column merge (:160-161) are spread over the warp's
// lanes, which cannot change the result.  Quirks that are part of the observable behaviour are
// reproduced: rows hold cids as floats, the merge test is `> 0` (cid 0 counts as absent), a
// connection matching three or more rows is dropped, limb 18 never starts a person, and peak
// scores are looked up by cid in the part-sorted table.
#include "common.cuh"

namespace ekp {

__global__ void __launch_bounds__(32) assemble_kernel(const ekp_peak* __restrict__ line, int max_peaks,
                                                      const Conn* __restrict__ conns, const int* __restrict__ n_conns,
                                                      int max_humans, float* __restrict__ subset_out,
                                                      int* __restrict__ num_humans, ekp_peak* __restrict__ hparts,
                                                      float* __restrict__ hscore, unsigned* __restrict__ overflow) {
    extern __shared__ float rows[];  // [max_humans][20]
    const int img = blockIdx.x, lane = threadIdx.x;
    const ekp_peak* L = line + (size_t) img * max_peaks;
    int nrows = 0;
    bool ovf = false;

    for (int limb = 0; limb < EKP_NUM_LIMB; limb++) {
        const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
        const int nc = min(n_conns[(size_t) img * EKP_NUM_LIMB + limb], EKP_MAX_PART);
        const Conn* C = conns + ((size_t) img * EKP_NUM_LIMB + limb) * EKP_MAX_PART;
        for (int k = 0; k < nc; k++) {
            const Conn cn = C[k];
            const float f1 = (float) cn.cid1, f2 = (float) cn.cid2;
            int found = 0, s1 = 0, s2 = 0;
            for (int base = 0; base < nrows; base += 32) {
                const int r = base + lane;
                const bool m = r < nrows && (rows[r * 20 + p1] == f1 || rows[r * 20 + p2] == f2);
                unsigned mask = __ballot_sync(0xffffffffu, m);
                const int cnt = __popc(mask);
                if (cnt) {
                    if (found == 0) {
                        s1 = base + __ffs(mask) - 1;
                        mask &= mask - 1;
                        if (mask) s2 = base + __ffs(mask) - 1;
                    } else if (found == 1) {
                        s2 = base + __ffs(mask) - 1;
                    }
                    found += cnt;
                }
            }
            if (found == 1) {
                if (lane == 0 && rows[s1 * 20 + p2] != f2) {
                    rows[s1 * 20 + p2] = f2;
                    rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], 1.0f);
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], __fadd_rn(L[cn.cid2].score, cn.score));
                }
            } else if (found == 2) {
                const bool both = lane < 18 && rows[s1 * 20 + lane] > 0.f && rows[s2 * 20 + lane] > 0.f;
                const bool membership = __any_sync(0xffffffffu, both);
                if (!membership) {
                    if (lane < 18) rows[s1 * 20 + lane] = __fadd_rn(rows[s1 * 20 + lane], __fadd_rn(rows[s2 * 20 + lane], 1.0f));
                    if (lane == 19) rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], rows[s2 * 20 + 19]);
                    if (lane == 18) {
                        float v = __fadd_rn(rows[s1 * 20 + 18], rows[s2 * 20 + 18]);
                        rows[s1 * 20 + 18] = __fadd_rn(v, cn.score);
                    }
                    __syncwarp();
                    if (lane < 20)  // erase row s2: every lane shifts its own column
                        for (int r = s2; r < nrows - 1; r++) rows[r * 20 + lane] = rows[(r + 1) * 20 + lane];
                    nrows--;
                } else if (lane == 0) {
                    rows[s1 * 20 + p2] = f2;
                    rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], 1.0f);
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], __fadd_rn(L[cn.cid2].score, cn.score));
                }
            } else if (found == 0 && limb < 18) {
                if (nrows < max_humans) {
                    if (lane < 20) {
                        float v = -1.0f;
                        if (lane == p1) v = f1;
                        if (lane == p2) v = f2;
                        if (lane == 19) v = 2.0f;
                        if (lane == 18) v = __fadd_rn(__fadd_rn(L[cn.cid1].score, L[cn.cid2].score), cn.score);
                        rows[nrows * 20 + lane] = v;
                    }
                    nrows++;
                } else {
                    ovf = true;
                }
            }
            __syncwarp();
        }
    }

    // prune (pafprocess.cpp:187-191): a reverse erase loop == an order-preserving filter
    int kept = 0;
    float* so = subset_out + (size_t) img * max_humans * 20;
    ekp_peak* hp = hparts + (size_t) img * max_humans * EKP_NUM_PART;
    for (int r = 0; r < nrows; r++) {
        const float cnt = rows[r * 20 + 19], sc = rows[r * 20 + 18];
        if (cnt < 4.0f || __fdiv_rn(sc, cnt) < 0.3f) continue;
        if (lane < 20) so[kept * 20 + lane] = rows[r * 20 + lane];
        if (lane < EKP_NUM_PART) {
            const int cid = (int) rows[r * 20 + lane];  // get_part_cid: float -> int
            ekp_peak o;
            if (cid >= 0) { const ekp_peak pk = L[cid]; o.x = pk.x; o.y = pk.y; o.score = pk.score; o.id = cid; }
            else { o.x = 0; o.y = 0; o.score = 0.f; o.id = -1; }
            hp[kept * EKP_NUM_PART + lane] = o;
        }
        if (lane == 0) hscore[(size_t) img * max_humans + kept] = __fdiv_rn(sc, cnt);  // get_score
        kept++;
    }
    if (lane == 0) {
        num_humans[img] = kept;
        if (ovf) atomicOr(overflow + img, EKP_OVF_HUMANS);
    }
}

cudaError_t configure_assemble(int max_humans) {
    return cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int) (sizeof(float) * 20 * (size_t) max_humans));
}

cudaError_t launch_assemble(const ekp_peak* line, int max_peaks, const Conn* conns, const int* n_conns, int max_humans,
                            int n, float* subset_out, int* num_humans, ekp_peak* hparts, float* hscore,
                            unsigned* overflow, cudaStream_t stream) {
    const size_t smem = sizeof(float) * 20 * (size_t) max_humans;
    assemble_kernel<<<n, 32, smem, stream>>>(line, max_peaks, conns, n_conns, max_humans, subset_out, num_humans, hparts,
                                             hscore, overflow);
    return cudaGetLastError();
}

}  // namespace ekp
