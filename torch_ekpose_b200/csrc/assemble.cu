// assemble.cu -- second half of stage 5: person assembly, pruning and the result record, one
// warp per image.  Replaces /root/reference/lib/pafprocess/pafprocess.cpp:127-191 (subset
// assembly and pruning) and the getter loop of paf_to_pose_cpp (paf_to_pose.py:361-377).
//
// The reference walks (limb, connection) strictly in order (pafprocess.cpp:130-185).  Within ONE limb,
// however, the connections carry distinct cid1 and distinct cid2 (the greedy step uses every peak at most
// once per limb side), so they can only interact through a subset row that two of them match, or through
// a merge.  Per limb, one lane per connection searches the rows as they stand at the start of the limb
// (pafprocess.cpp:137-144); if every connection matches at most one row and no row is matched twice, all
// of them are applied at once -- extend (:146-151) or start a row (:173-183, new rows numbered in
// connection order) -- which is exactly what the sequential walk produces (argument at limb_parallel).
// Otherwise (a merge, a shared row: rare) the limb is walked sequentially, rows searched by the lanes.
// Quirks that are part of the observable behaviour are reproduced: rows hold cids as floats, the merge
// test is `> 0` (cid 0 counts as absent), a connection matching three or more rows is dropped, limb 18
// never starts a person, and peak scores are looked up by cid in the part-sorted table.
//
// One warp per image; the image's connections and peak scores are staged in shared memory first
// (coalesced, one flat pass); results go to one packed record per image (ResultLayout): one
// device-to-host copy per batch.
#include "common.cuh"

namespace ekp {

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

__device__ __forceinline__ ConnRec make_rec(const Conn& cn) {
    ConnRec r;
    r.f1 = (float) cn.cid1;
    r.f2 = (float) cn.cid2;
    r.score = cn.score;
    r.s_ext = cn.s_ext;  // formed by paf_connect_kernel
    r.s_new = cn.s_new;
    r.pad0 = r.pad1 = r.pad2 = 0.f;
    return r;
}

struct AsmInput {
    const ConnRec* sRec;  // staged records (or nullptr -> build from `conns` on the fly)
    const int* sStart;    // [20] prefix of per-limb counts
    const Conn* conns;    // this image's [19][EKP_MAX_PART]
    __device__ __forceinline__ ConnRec rec_at(int limb, int k) const {
        return sRec ? sRec[sStart[limb] + k] : make_rec(conns[(size_t) limb * EKP_MAX_PART + k]);
    }
};

// (forward: argument for the parallel application is given at limb_parallel below)
// K connections per lane (k = lane + 32 q), everything in registers: the usual case.
template <int K>
__device__ __forceinline__ bool limb_parallel_regs(const AsmInput& in, int limb, int max_humans, float* __restrict__ rows,
                                                   int& nrows_io, bool& ovf, int* __restrict__ sClaim) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
    const int nc = in.sStart[limb + 1] - in.sStart[limb];
    const int nrows = nrows_io;
    ConnRec cn[K];
    int found[K], s1[K];
    bool active[K];
#pragma unroll
    for (int q = 0; q < K; q++) {
        active[q] = lane + 32 * q < nc;
        found[q] = 0; s1[q] = -1;
        if (active[q]) cn[q] = in.rec_at(limb, lane + 32 * q);
    }
    for (int r = 0; r < nrows; r++) {  // every lane reads the same two words: broadcast
        const float a = rows[r * 20 + p1], b = rows[r * 20 + p2];
#pragma unroll
        for (int q = 0; q < K; q++)
            if (active[q] && (a == cn[q].f1 || b == cn[q].f2)) { found[q]++; s1[q] = r; }
    }
    bool bad = false;
    if (K == 1) {  // a row matched by two connections shows up as two lanes with the same s1
        const unsigned peers = __match_any_sync(FULL, s1[0] >= 0 ? s1[0] : -1 - lane);
        bad = found[0] >= 2 || (s1[0] >= 0 && (peers & (peers - 1)) != 0u);
    } else {       // ... or as a claim that does not read back
#pragma unroll
        for (int q = 0; q < K; q++)
            if (s1[q] >= 0) sClaim[s1[q]] = lane + 32 * q;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < K; q++) bad |= found[q] >= 2 || (s1[q] >= 0 && sClaim[s1[q]] != lane + 32 * q);
    }
    if (__any_sync(FULL, bad)) return false;
    int nnew = 0;
#pragma unroll
    for (int q = 0; q < K; q++) {
        if (s1[q] >= 0) {  // found == 1, pafprocess.cpp:146-151
            if (rows[s1[q] * 20 + p2] != cn[q].f2) {
                rows[s1[q] * 20 + p2] = cn[q].f2;
                rows[s1[q] * 20 + 19] = __fadd_rn(rows[s1[q] * 20 + 19], 1.0f);
                rows[s1[q] * 20 + 18] = __fadd_rn(rows[s1[q] * 20 + 18], cn[q].s_ext);
            }
        }
        const bool starts = active[q] && s1[q] < 0 && limb < 18;  // found == 0, :173-183
        const unsigned mask = __ballot_sync(FULL, starts);
        if (starts) {
            const int r = nrows + nnew + __popc(mask & ((1u << lane) - 1u));
            if (r < max_humans) {
#pragma unroll
                for (int c = 0; c < 18; c++) rows[r * 20 + c] = -1.0f;
                rows[r * 20 + p1] = cn[q].f1;
                rows[r * 20 + p2] = cn[q].f2;
                rows[r * 20 + 18] = cn[q].s_new;
                rows[r * 20 + 19] = 2.0f;
            }
        }
        nnew += __popc(mask);
    }
    if (nrows + nnew > max_humans) { ovf = true; nnew = max_humans - nrows; }
    nrows_io = nrows + nnew;
    __syncwarp();
    return true;
}

// ---- one limb, all connections at once -------------------------------------------------------------
// Let the limb's connections be k = 0..nc-1 (acceptance order), with distinct f1 (cid1) and distinct f2
// (cid2).  Against the rows at the START of the limb, connection k matches found_k rows.  Claim: if every
// found_k <= 1 and the matched rows are pairwise different, the sequential walk takes, for every k, the same
// branch on the same row as it would at the start of the limb:
//  * an earlier j that EXTENDS its row R_j only changes R_j[p2] (to f2_j != f2_k), R_j[18], R_j[19]; k did
//    not match R_j, and cannot start to (R_j[p1] is unchanged, R_j[p2] becomes f2_j != f2_k);
//  * an earlier j that STARTS a row gives it p1 = f1_j != f1_k and p2 = f2_j != f2_k: no match for k;
//  * a j whose single match is through p2 only changes nothing (pafprocess.cpp:147).
// So extensions touch pairwise different rows, new rows are appended in connection order, and the result
// equals the sequential one.  Returns false (state untouched) when the condition does not hold.
__device__ bool limb_parallel(const AsmInput& in, int limb, int max_humans, float* __restrict__ rows, int& nrows_io, bool& ovf,
                              int* __restrict__ sClaim /* [max_humans] */, short* __restrict__ sMatch /* [EKP_MAX_PART] */) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
    const int nc = in.sStart[limb + 1] - in.sStart[limb];
    const int nrows = nrows_io;
    if (nc <= 32) return limb_parallel_regs<1>(in, limb, max_humans, rows, nrows_io, ovf, sClaim);
    if (nc <= 64) return limb_parallel_regs<2>(in, limb, max_humans, rows, nrows_io, ovf, sClaim);
    // search: matched row (or -1: none) per connection; two or more matches end the attempt
    bool bad = false;
    for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        if (k < nc) {
            const ConnRec cn = in.rec_at(limb, k);
            int found = 0, s1 = -1;
            for (int r = 0; r < nrows; r++) {  // every lane reads the same two words: broadcast
                const bool m = rows[r * 20 + p1] == cn.f1 || rows[r * 20 + p2] == cn.f2;
                if (m) { found++; s1 = r; }
            }
            if (found >= 2) bad = true;
            sMatch[k] = (short) s1;
        }
    }
    if (__any_sync(FULL, bad)) return false;
    __syncwarp();
    for (int k = lane; k < nc; k += 32)  // no row may be matched by two connections
        if (sMatch[k] >= 0) sClaim[sMatch[k]] = k;
    __syncwarp();
    for (int k = lane; k < nc; k += 32)
        if (sMatch[k] >= 0 && sClaim[sMatch[k]] != k) bad = true;
    if (__any_sync(FULL, bad)) return false;
    // apply
    int nnew = 0;  // rows started so far by this limb (identical in every lane)
    for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        bool starts = false;
        ConnRec cn;
        if (k < nc) {
            cn = in.rec_at(limb, k);
            const int s1 = sMatch[k];
            if (s1 >= 0) {  // found == 1, pafprocess.cpp:146-151
                if (rows[s1 * 20 + p2] != cn.f2) {
                    rows[s1 * 20 + p2] = cn.f2;
                    rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], 1.0f);
                    rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], cn.s_ext);
                }
            } else {
                starts = limb < 18;  // found == 0, :173-183
            }
        }
        const unsigned mask = __ballot_sync(FULL, starts);
        if (starts) {
            const int r = nrows + nnew + __popc(mask & ((1u << lane) - 1u));
            if (r < max_humans) {
#pragma unroll
                for (int q = 0; q < 18; q++) rows[r * 20 + q] = -1.0f;
                rows[r * 20 + p1] = cn.f1;
                rows[r * 20 + p2] = cn.f2;
                rows[r * 20 + 18] = cn.s_new;
                rows[r * 20 + 19] = 2.0f;
            }
        }
        nnew += __popc(mask);
    }
    if (nrows + nnew > max_humans) { ovf = true; nnew = max_humans - nrows; }
    nrows_io = nrows + nnew;
    __syncwarp();
    return true;
}

// ---- one limb, connection by connection (the reference's walk) ---------------------------------------
// The row search is spread over the lanes; for up to 128 rows the two columns it compares (p1, p2) live in
// registers (row r in lane r & 31, slot r >> 5) and are kept in step with the rows in shared memory, so a step
// costs ballots instead of shared-memory round trips.
constexpr int kSeqSlots = 4;
__device__ void limb_sequential(const AsmInput& in, int limb, int max_humans, float* __restrict__ rows, int& nrows_io, bool& ovf) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    int nrows = nrows_io;
    const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
    const int nc = in.sStart[limb + 1] - in.sStart[limb];
    const bool cached = max_humans <= 32 * kSeqSlots;
    float c1[kSeqSlots], c2[kSeqSlots];
    auto reload = [&]() {
#pragma unroll
        for (int s = 0; s < kSeqSlots; s++) {
            const int r = 32 * s + lane;
            c1[s] = r < nrows ? rows[r * 20 + p1] : -1.0f;
            c2[s] = r < nrows ? rows[r * 20 + p2] : -1.0f;
        }
    };
    if (cached) reload();
    for (int k = 0; k < nc; k++) {
        const ConnRec cn = in.rec_at(limb, k);
        const float f1 = cn.f1, f2 = cn.f2;
        int found = 0, s1 = 0, s2 = 0;
        if (cached) {
#pragma unroll
            for (int s = 0; s < kSeqSlots; s++) {
                if (32 * s >= nrows) break;
                unsigned mask = __ballot_sync(FULL, 32 * s + lane < nrows && (c1[s] == f1 || c2[s] == f2));
                const int c = __popc(mask);
                if (c) {
                    if (found == 0) {
                        s1 = 32 * s + __ffs(mask) - 1;
                        mask &= mask - 1;
                        if (mask) s2 = 32 * s + __ffs(mask) - 1;
                    } else if (found == 1) {
                        s2 = 32 * s + __ffs(mask) - 1;
                    }
                    found += c;
                }
            }
        } else {
            for (int base = 0; base < nrows; base += 32) {
                const int r = base + lane;
                const bool m = r < nrows && (rows[r * 20 + p1] == f1 || rows[r * 20 + p2] == f2);
                unsigned mask = __ballot_sync(FULL, m);
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
        }
        auto extend_s1 = [&]() {  // rows[s1][p2] = cid2, count + 1, score + (peak(cid2) + conn)
            if (lane == 0) {
                rows[s1 * 20 + p2] = f2;
                rows[s1 * 20 + 19] = __fadd_rn(rows[s1 * 20 + 19], 1.0f);
                rows[s1 * 20 + 18] = __fadd_rn(rows[s1 * 20 + 18], cn.s_ext);
            }
            if (cached && lane == (s1 & 31)) {
#pragma unroll
                for (int s = 0; s < kSeqSlots; s++)
                    if (s == (s1 >> 5)) c2[s] = f2;
            }
        };
        if (found == 1) {
            // rows[s1][p2] != cid2, read from the cache's owner or from shared memory
            const bool differs = rows[s1 * 20 + p2] != f2;
            if (differs) extend_s1();
        } else if (found == 2) {
            const bool both = lane < 18 && rows[s1 * 20 + lane] > 0.f && rows[s2 * 20 + lane] > 0.f;
            const bool membership = __any_sync(FULL, both);
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
                __syncwarp();
                if (cached) reload();  // rare: rows moved
            } else {
                extend_s1();
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
                if (cached && lane == (nrows & 31)) {
#pragma unroll
                    for (int s = 0; s < kSeqSlots; s++)
                        if (s == (nrows >> 5)) { c1[s] = f1; c2[s] = f2; }
                }
                nrows++;
            } else {
                ovf = true;
            }
        }
        __syncwarp();
    }
    nrows_io = nrows;
}

#ifdef EKP_ASM_PROFILE  // tools/ only: time per phase summed over images (ns) and limb counts
__device__ unsigned long long g_asm_prof[8];
__device__ __forceinline__ unsigned long long asm_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define APROF(k) do { if (threadIdx.x == 0) { const unsigned long long _t = asm_now(); atomicAdd(&g_asm_prof[k], _t - prof_t); prof_t = _t; } } while (0)
#define ACOUNT(k) do { if (threadIdx.x == 0) { atomicAdd(&g_asm_prof[k], 1ull); if (k == 5) prof_nseq++; } } while (0)
extern "C" int ekp_debug_asm_profile(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, g_asm_prof, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_asm_prof, z, sizeof(z)); }
    return 0;
}
#else
#define APROF(k) do { } while (0)
#define ACOUNT(k) do { } while (0)
#endif

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
    __shared__ short sMatch[EKP_MAX_PART];  // limb_parallel: matched row per connection

    const int img = blockIdx.x, lane = threadIdx.x;
#ifdef EKP_ASM_PROFILE
    unsigned long long prof_t = asm_now();
    const unsigned long long prof_t0 = prof_t;
    unsigned long long prof_nseq = 0;
#endif
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
        int start[EKP_NUM_LIMB];  // per-limb offsets in registers: the limb of a flat index costs no memory access
#pragma unroll
        for (int l = 0; l < EKP_NUM_LIMB; l++) start[l] = sStart[l];
#pragma unroll 4
        for (int idx = lane; idx < total_conns; idx += 32) {
            int limb = 0, base = 0;
#pragma unroll
            for (int l = 1; l < EKP_NUM_LIMB; l++)
                if (idx >= start[l]) { limb = l; base = start[l]; }
            sRec[idx] = make_rec(Cimg[(size_t) limb * EKP_MAX_PART + (idx - base)]);
        }
    }
    __syncwarp();
    AsmInput in;
    in.sRec = staged ? sRec : nullptr;
    in.sStart = sStart;
    in.conns = Cimg;

    APROF(0);  // staging
    // ---- assembly, limb by limb (pafprocess.cpp:130-185) ------------------------------------------
    int nrows = 0;
    bool ovf = false;
    for (int limb = 0; limb < EKP_NUM_LIMB; limb++) {
        if (sStart[limb + 1] == sStart[limb]) continue;
        if (!limb_parallel(in, limb, max_humans, rows, nrows, ovf, sKept /* free until the prune */, sMatch)) {
            APROF(1);
            limb_sequential(in, limb, max_humans, rows, nrows, ovf);
            APROF(2); ACOUNT(5);
        } else { APROF(1); ACOUNT(4); }
        __syncwarp();
    }

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
    APROF(3);  // prune + record
#ifdef EKP_ASM_PROFILE
    if (threadIdx.x == 0) { atomicMax(&g_asm_prof[6], asm_now() - prof_t0); atomicMax(&g_asm_prof[7], prof_nseq); }
#endif
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
    return raise_dynamic_smem_limit(assemble_kernel, assemble_smem(max_humans, assemble_conn_cap(max_peaks)));
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
