// assemble.cu -- second half of stage 5: person assembly, pruning and the result record, one
// warp per image.  Replaces /root/reference/lib/pafprocess/pafprocess.cpp:127-191 (subset
// assembly and pruning) and the getter loop of paf_to_pose_cpp (paf_to_pose.py:361-377).
//
// The assembly is inherently sequential over (limb, connection) and is kept so; only the row
// SEARCH (pafprocess.cpp:137-144) and the 18-column merge (:160-161) are spread over the warp's
// lanes, which cannot change the result.  Quirks that are part of the observable behaviour are
// reproduced: rows hold cids as floats, the merge test is `> 0` (cid 0 counts as absent), a
// connection matching three or more rows is dropped, limb 18 never starts a person, and peak
// scores are looked up by cid in the part-sorted table.
//
// The loop is latency bound, so everything it touches (the image's connections and peak scores)
// is first staged in shared memory with coalesced loads; the serial part then only sees ~30-cycle
// shared-memory latencies instead of dependent global loads.  Results go to one packed record per
// image (ResultLayout) so the host needs a single device-to-host copy per batch.
#include "common.cuh"

namespace ekp {

__global__ void __launch_bounds__(32) assemble_kernel(const ekp_peak* __restrict__ line, int max_peaks,
                                                      const int* __restrict__ n_peaks, const Conn* __restrict__ conns,
                                                      const int* __restrict__ n_conns, int max_humans, int conn_cap,
                                                      int score_cap, const unsigned* __restrict__ overflow,
                                                      unsigned char* __restrict__ records, ResultLayout lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);                        // [max_humans][20]
    Conn* sConn = reinterpret_cast<Conn*>(rows + (size_t) max_humans * 20);   // [conn_cap]
    float* sScore = reinterpret_cast<float*>(sConn + conn_cap);              // [score_cap]
    __shared__ int sStart[EKP_NUM_LIMB + 1];

    const int img = blockIdx.x, lane = threadIdx.x;
    const ekp_peak* L = line + (size_t) img * max_peaks;
    const int npk = n_peaks[img];

    // ---- stage connections and peak scores --------------------------------------------------
    int cnt = 0;
    if (lane < EKP_NUM_LIMB) cnt = min(n_conns[(size_t) img * EKP_NUM_LIMB + lane], EKP_MAX_PART);
    int incl = cnt;  // inclusive prefix over the 19 limbs
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane < EKP_NUM_LIMB) sStart[lane] = incl - cnt;
    if (lane == EKP_NUM_LIMB - 1) sStart[EKP_NUM_LIMB] = incl;
    __syncwarp();
    const int total_conns = sStart[EKP_NUM_LIMB];
    const bool staged = total_conns <= conn_cap && npk <= score_cap;
    if (staged) {
        for (int limb = 0; limb < EKP_NUM_LIMB; limb++) {
            const int s0 = sStart[limb], n = sStart[limb + 1] - s0;
            const Conn* C = conns + ((size_t) img * EKP_NUM_LIMB + limb) * EKP_MAX_PART;
            for (int k = lane; k < n; k += 32) sConn[s0 + k] = C[k];
        }
        for (int k = lane; k < npk; k += 32) sScore[k] = L[k].score;
    }
    __syncwarp();

    auto conn_at = [&](int limb, int k) -> Conn {
        return staged ? sConn[sStart[limb] + k] : conns[((size_t) img * EKP_NUM_LIMB + limb) * EKP_MAX_PART + k];
    };
    auto score_of = [&](int cid) -> float { return staged ? sScore[cid] : L[cid].score; };

    // ---- sequential assembly ------------------------------------------------------------------
    int nrows = 0;
    bool ovf = false;
    for (int limb = 0; limb < EKP_NUM_LIMB; limb++) {
        const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
        const int nc = sStart[limb + 1] - sStart[limb];
        for (int k = 0; k < nc; k++) {
            const Conn cn = conn_at(limb, k);
            const float f1 = (float) cn.cid1, f2 = (float) cn.cid2;
            int found = 0, s1 = 0, s2 = 0;
            for (int base = 0; base < nrows; base += 32) {
                const int r = base + lane;
                const bool m = r < nrows && (rows[r * 20 + p1] == f1 || rows[r * 20 + p2] == f2);
                unsigned mask = __ballot_sync(0xffffffffu, m);
                const int c = __popc(mask);
                if (c) {
                    if (found == 0) {
                        s1 = base + __ffs(mask) - 1;
                        mask &= mask - 1;
                        if (mask) s2 = base + __ffs(mask) - 1;
                    } else if (found == 1) {
                        s2 = base + __ffs(mask) - 1;
                    }
                    found += c;
                }
            }
            if (found == 1) {
                if (lane == 0 && rows[s1 * 20 + p2] != f2) {
                    rows[s1 * 20 + p2] = f2;
                    rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], 1.0f);
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], __fadd_rn(score_of(cn.cid2), cn.score));
                }
            } else if (found == 2) {
                const bool both = lane < 18 && rows[s1 * 20 + lane] > 0.f && rows[s2 * 20 + lane] > 0.f;
                const bool membership = __any_sync(0xffffffffu, both);
                if (!membership) {
                    if (lane < 18) rows[s1 * 20 + lane] = __fadd_rn(rows[s1 * 20 + lane], __fadd_rn(rows[s2 * 20 + lane], 1.0f));
                    if (lane == 19) rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], rows[s2 * 20 + 19]);
                    if (lane == 18) {
                        const float v = __fadd_rn(rows[s1 * 20 + 18], rows[s2 * 20 + 18]);
                        rows[s1 * 20 + 18] = __fadd_rn(v, cn.score);
                    }
                    __syncwarp();
                    if (lane < 20)  // erase row s2: every lane shifts its own column
                        for (int r = s2; r < nrows - 1; r++) rows[r * 20 + lane] = rows[(r + 1) * 20 + lane];
                    nrows--;
                } else if (lane == 0) {
                    rows[s1 * 20 + p2] = f2;
                    rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], 1.0f);
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], __fadd_rn(score_of(cn.cid2), cn.score));
                }
            } else if (found == 0 && limb < 18) {
                if (nrows < max_humans) {
                    if (lane < 20) {
                        float v = -1.0f;
                        if (lane == p1) v = f1;
                        if (lane == p2) v = f2;
                        if (lane == 19) v = 2.0f;
                        if (lane == 18) v = __fadd_rn(__fadd_rn(score_of(cn.cid1), score_of(cn.cid2)), cn.score);
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

    // ---- prune (pafprocess.cpp:187-191: a reverse erase loop == an order-preserving filter) and
    //      write the image's result record ------------------------------------------------------
    unsigned char* rec = records + (size_t) img * lay.stride;
    float* so = reinterpret_cast<float*>(rec + lay.off_subset);
    ekp_peak* hp = reinterpret_cast<ekp_peak*>(rec + lay.off_hparts);
    float* hs = reinterpret_cast<float*>(rec + lay.off_hscore);
    int kept = 0;
    for (int r = 0; r < nrows; r++) {
        const float c = rows[r * 20 + 19], sc = rows[r * 20 + 18];
        if (c < 4.0f || __fdiv_rn(sc, c) < 0.3f) continue;
        if (lane < 20) so[kept * 20 + lane] = rows[r * 20 + lane];
        if (lane < EKP_NUM_PART) {
            const int cid = (int) rows[r * 20 + lane];  // get_part_cid: float -> int
            ekp_peak o;
            if (cid >= 0) { const ekp_peak pk = L[cid]; o.x = pk.x; o.y = pk.y; o.score = pk.score; o.id = cid; }
            else { o.x = 0; o.y = 0; o.score = 0.f; o.id = -1; }
            hp[kept * EKP_NUM_PART + lane] = o;
        }
        if (lane == 0) hs[kept] = __fdiv_rn(sc, c);  // get_score
        kept++;
    }
    if (lane == 0) {
        int4 head;
        head.x = kept;
        head.y = npk;
        head.z = (int) (overflow[img] | (ovf ? EKP_OVF_HUMANS : 0u));
        head.w = 0;
        *reinterpret_cast<int4*>(rec) = head;
    }
}

static size_t assemble_smem(int max_humans, int conn_cap, int score_cap) {
    return sizeof(float) * 20 * (size_t) max_humans + sizeof(Conn) * (size_t) conn_cap + sizeof(float) * (size_t) score_cap;
}
void assemble_caps(int max_peaks, int* conn_cap, int* score_cap) {
    *conn_cap = 2 * max_peaks < 2048 ? 2 * max_peaks : 2048;
    *score_cap = max_peaks < 4096 ? max_peaks : 4096;
}

cudaError_t configure_assemble(int max_humans, int max_peaks) {
    int cc, sc;
    assemble_caps(max_peaks, &cc, &sc);
    return cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) assemble_smem(max_humans, cc, sc));
}

cudaError_t launch_assemble(const ekp_peak* line, int max_peaks, const int* n_peaks, const Conn* conns, const int* n_conns,
                            int max_humans, int n, const unsigned* overflow, unsigned char* records, const ResultLayout& lay,
                            cudaStream_t stream) {
    int cc, sc;
    assemble_caps(max_peaks, &cc, &sc);
    assemble_kernel<<<n, 32, assemble_smem(max_humans, cc, sc), stream>>>(line, max_peaks, n_peaks, conns, n_conns, max_humans, cc,
                                                                        sc, overflow, records, lay);
    return cudaGetLastError();
}

}  // namespace ekp
