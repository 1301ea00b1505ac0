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
// A single warp running dependent code pays full latency on every instruction, so:
//  * the image's connections and peak scores are staged in shared memory first (coalesced);
//  * up to 64 candidate people live in REGISTERS, two subset rows per lane (the 19-limb loop is
//    unrolled so every column index is a compile-time constant); extending or starting a row costs
//    one broadcast load, two compares and two ballots.  Merges and a 65th row continue, from the
//    same connection, on the general shared-memory path below (same arithmetic, any row count);
//  * results go to one packed record per image (ResultLayout): one device-to-host copy per batch.
#include "common.cuh"

namespace ekp {

// pafprocess.h:21-24 as compile-time constants for the unrolled fast path
__host__ __device__ constexpr int limb_a(int l) {
    constexpr int t[EKP_NUM_LIMB] = {1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5};
    return t[l];
}
__host__ __device__ constexpr int limb_b(int l) {
    constexpr int t[EKP_NUM_LIMB] = {2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17};
    return t[l];
}

// Everything the sequential loop needs about one connection, prepared in parallel while staging:
// the cids as floats (rows hold floats) and the two score sums the reference forms
// (pafprocess.cpp:150/171: peak(cid2) + conn;  :179-181: (peak(cid1) + peak(cid2)) + conn), same order.
struct __align__(16) ConnRec {
    float f1, f2;   // (float) cid1, (float) cid2
    float score;    // connection score
    float s_ext;    // peak_score(cid2) + score        (a row is extended by part2)
    float s_new;    // (peak_score(cid1) + peak_score(cid2)) + score   (a new row)
    float pad0, pad1, pad2;
};

__device__ __forceinline__ ConnRec make_rec(const Conn& cn, const ekp_peak* L) {
    ConnRec r;
    r.f1 = (float) cn.cid1;
    r.f2 = (float) cn.cid2;
    r.score = cn.score;
    const float p1 = L[cn.cid1].score, p2 = L[cn.cid2].score;
    r.s_ext = __fadd_rn(p2, cn.score);
    r.s_new = __fadd_rn(__fadd_rn(p1, p2), cn.score);
    r.pad0 = r.pad1 = r.pad2 = 0.f;
    return r;
}

struct AsmInput {
    const ConnRec* sRec;  // staged records (or nullptr -> build from `conns` / `L` on the fly)
    const int* sStart;    // [20] prefix of per-limb counts
    const Conn* conns;    // this image's [19][EKP_MAX_PART]
    const ekp_peak* L;    // this image's part-sorted peak table
    __device__ __forceinline__ ConnRec rec_at(int limb, int k) const {
        return sRec ? sRec[sStart[limb] + k] : make_rec(conns[(size_t) limb * EKP_MAX_PART + k], L);
    }
};

// ---- fast path: R subset rows per lane, in registers ------------------------------------------
// Row index i lives in lane (i & 31), slot (i >> 5); capacity 32*R rows.  It handles the two cases
// that make up almost every step -- a connection extends one row (found == 1) or starts a new one
// (found == 0) -- with one broadcast load, R compares and R ballots.  When a connection matches
// TWO rows (a merge, rare) or a row beyond the capacity is needed, the rows are written to shared
// memory and the caller continues from exactly that connection on the general path.  The 19-limb
// loop is unrolled (column indices must be compile-time for registers), so the per-limb body is
// kept this small on purpose: the whole path has to stay inside the instruction cache.
template <int R>
__device__ void assemble_in_registers(const AsmInput& in, int max_humans, float* __restrict__ rows_out, int& nrows_out,
                                      int& resume_limb, int& resume_k) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    float r[R][20];
#pragma unroll
    for (int s = 0; s < R; s++)
#pragma unroll
        for (int q = 0; q < 20; q++) r[s][q] = -1.0f;
    int nrows = 0;
    const int cap = max_humans < 32 * R ? max_humans : 32 * R;
    resume_limb = EKP_NUM_LIMB;  // "finished"
    resume_k = 0;
    bool bail = false;
#pragma unroll
    for (int limb = 0; limb < EKP_NUM_LIMB; limb++) {
        const int p1 = limb_a(limb), p2 = limb_b(limb);
        const int nc = in.sStart[limb + 1] - in.sStart[limb];
        for (int k = 0; k < nc; k++) {
            const ConnRec cn = in.rec_at(limb, k);
            int found = 0;
            bool m[R];
#pragma unroll
            for (int s = 0; s < R; s++) {  // row search, pafprocess.cpp:137-144
                m[s] = (32 * s + lane) < nrows && (r[s][p1] == cn.f1 || r[s][p2] == cn.f2);
                found += __popc(__ballot_sync(FULL, m[s]));
            }
            if (found == 1) {
#pragma unroll
                for (int s = 0; s < R; s++)
                    if (m[s] && r[s][p2] != cn.f2) {
                        r[s][p2] = cn.f2;
                        r[s][19] = __fadd_rn(r[s][19], 1.0f);
                        r[s][18] = __fadd_rn(r[s][18], cn.s_ext);
                    }
            } else if (found == 0) {
                if (limb < 18) {
                    if (nrows >= cap) { bail = true; resume_limb = limb; resume_k = k; break; }
#pragma unroll
                    for (int s = 0; s < R; s++)
                        if (32 * s + lane == nrows) {
#pragma unroll
                            for (int q = 0; q < 18; q++) r[s][q] = -1.0f;
                            r[s][p1] = cn.f1;
                            r[s][p2] = cn.f2;
                            r[s][19] = 2.0f;
                            r[s][18] = cn.s_new;
                        }
                    nrows++;
                }
            } else if (found == 2) {  // merge or extend-with-conflict: continue on the general path
                bail = true; resume_limb = limb; resume_k = k;
                break;
            }  // found >= 3: the reference takes no branch
        }
        if (bail) break;
    }
#pragma unroll
    for (int s = 0; s < R; s++)
        if (32 * s + lane < nrows) {
#pragma unroll
            for (int q = 0; q < 20; q++) rows_out[(32 * s + lane) * 20 + q] = r[s][q];
        }
    nrows_out = nrows;
    __syncwarp();
}

// ---- general path: rows in shared memory, any count up to max_humans ---------------------------
// Starts at connection k0 of limb limb0 with nrows_io rows already in `rows`.
__device__ void assemble_in_smem(const AsmInput& in, int max_humans, float* __restrict__ rows, int& nrows_io, bool& ovf,
                                 int limb0, int k0) {
    const int lane = threadIdx.x;
    int nrows = nrows_io;
    for (int limb = limb0; limb < EKP_NUM_LIMB; limb++) {
        const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
        const int nc = in.sStart[limb + 1] - in.sStart[limb];
        for (int k = (limb == limb0 ? k0 : 0); k < nc; k++) {
            const ConnRec cn = in.rec_at(limb, k);
            const float f1 = cn.f1, f2 = cn.f2;
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
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], cn.s_ext);
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
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], cn.s_ext);
                }
            } else if (found == 0 && limb < 18) {
                if (nrows < max_humans) {
                    if (lane < 20) {
                        float v = -1.0f;
                        if (lane == p1) v = f1;
                        if (lane == p2) v = f2;
                        if (lane == 19) v = 2.0f;
                        if (lane == 18) v = cn.s_new;
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
    nrows_io = nrows;
}

__global__ void __launch_bounds__(32) assemble_kernel(const ekp_peak* __restrict__ line, int max_peaks,
                                                      const int* __restrict__ n_peaks, const Conn* __restrict__ conns,
                                                      const int* __restrict__ n_conns, int max_humans, int conn_cap,
                                                      const unsigned* __restrict__ overflow,
                                                      unsigned char* __restrict__ records, ResultLayout lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);                                             // [max(max_humans, 32)][20]
    ConnRec* sRec = reinterpret_cast<ConnRec*>(rows + (size_t) (max_humans < 32 ? 32 : max_humans) * 20);  // [conn_cap]
    int* sKept = reinterpret_cast<int*>(sRec + conn_cap);                                         // [max_humans]
    __shared__ int sStart[EKP_NUM_LIMB + 1];

    const int img = blockIdx.x, lane = threadIdx.x;
    const ekp_peak* L = line + (size_t) img * max_peaks;
    const Conn* Cimg = conns + (size_t) img * EKP_NUM_LIMB * EKP_MAX_PART;
    const int npk = n_peaks[img];

    // ---- stage one prepared record per connection ------------------------------------------------
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
    const bool staged = total_conns <= conn_cap;
    if (staged) {
        // one flat pass so that all loads are in flight together (a per-limb loop would pay one
        // global-memory round trip per limb)
        for (int idx = lane; idx < total_conns; idx += 32) {
            int limb = 0;
#pragma unroll
            for (int l = 1; l < EKP_NUM_LIMB; l++) limb += (idx >= sStart[l]);
            sRec[idx] = make_rec(Cimg[(size_t) limb * EKP_MAX_PART + (idx - sStart[limb])], L);
        }
    }
    __syncwarp();
    AsmInput in;
    in.sRec = staged ? sRec : nullptr;
    in.sStart = sStart;
    in.conns = Cimg;
    in.L = L;

    // ---- sequential assembly ------------------------------------------------------------------
    int nrows = 0;
    bool ovf = false;
    // Registers hold up to 64 rows (two per lane); whatever they cannot do (merges, more rows) continues
    // on the shared-memory path from the connection where they stopped.
    int resume_limb = 0, resume_k = 0;
    if (max_humans >= 1) assemble_in_registers<2>(in, max_humans, rows, nrows, resume_limb, resume_k);
    if (resume_limb < EKP_NUM_LIMB) assemble_in_smem(in, max_humans, rows, nrows, ovf, resume_limb, resume_k);
    __syncwarp();

    // ---- prune (pafprocess.cpp:187-191: a reverse erase loop == an order-preserving filter) and
    //      write the image's result record, all lanes busy --------------------------------------
    unsigned char* rec = records + (size_t) img * lay.stride;
    float* so = reinterpret_cast<float*>(rec + lay.off_subset);
    ekp_peak* hp = reinterpret_cast<ekp_peak*>(rec + lay.off_hparts);
    float* hs = reinterpret_cast<float*>(rec + lay.off_hscore);
    int kept = 0;
    for (int base = 0; base < nrows; base += 32) {
        const int r = base + lane;
        bool keep = false;
        if (r < nrows) {
            const float c = rows[r * 20 + 19], sc = rows[r * 20 + 18];
            keep = !(c < 4.0f || __fdiv_rn(sc, c) < 0.3f);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) sKept[kept + __popc(mask & ((1u << lane) - 1u))] = r;
        kept += __popc(mask);
    }
    __syncwarp();
    for (int idx = lane; idx < kept * 20; idx += 32) {
        const int k = idx / 20, q = idx - k * 20;
        so[idx] = rows[sKept[k] * 20 + q];
    }
    for (int idx = lane; idx < kept * EKP_NUM_PART; idx += 32) {
        const int k = idx / EKP_NUM_PART, q = idx - k * EKP_NUM_PART;
        const int cid = (int) rows[sKept[k] * 20 + q];  // get_part_cid: float -> int
        ekp_peak o;
        if (cid >= 0) { const ekp_peak pk = L[cid]; o.x = pk.x; o.y = pk.y; o.score = pk.score; o.id = cid; }
        else { o.x = 0; o.y = 0; o.score = 0.f; o.id = -1; }
        hp[idx] = o;
    }
    for (int k = lane; k < kept; k += 32) {
        const int r = sKept[k];
        hs[k] = __fdiv_rn(rows[r * 20 + 18], rows[r * 20 + 19]);  // get_score
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

static size_t assemble_smem(int max_humans, int conn_cap) {
    const size_t nrow = (size_t) (max_humans < 32 ? 32 : max_humans);
    return sizeof(float) * 20 * nrow + sizeof(ConnRec) * (size_t) conn_cap + sizeof(int) * nrow;
}
static int assemble_conn_cap(int max_peaks) { return 2 * max_peaks < 1536 ? 2 * max_peaks : 1536; }

cudaError_t configure_assemble(int max_humans, int max_peaks) {
    return cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int) assemble_smem(max_humans, assemble_conn_cap(max_peaks)));
}

cudaError_t launch_assemble(const ekp_peak* line, int max_peaks, const int* n_peaks, const Conn* conns, const int* n_conns,
                            int max_humans, int n, const unsigned* overflow, unsigned char* records, const ResultLayout& lay,
                            cudaStream_t stream) {
    const int cc = assemble_conn_cap(max_peaks);
    assemble_kernel<<<n, 32, assemble_smem(max_humans, cc), stream>>>(line, max_peaks, n_peaks, conns, n_conns, max_humans, cc,
                                                                    overflow, records, lay);
    return cudaGetLastError();
}

}  // namespace ekp
