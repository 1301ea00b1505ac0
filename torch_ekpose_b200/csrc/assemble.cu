// assemble.cu -- second half of stage 5: person assembly, pruning and the result record, one
// block per image.  Replaces /root/reference/lib/pafprocess/pafprocess.cpp:127-191 (subset
// assembly and pruning) and the getter loop of paf_to_pose_cpp (paf_to_pose.py:361-377).
//
// The reference walks (limb, connection) strictly in order (pafprocess.cpp:130-185) and, per connection, scans
// every subset row for `row[part1] == cid1 || row[part2] == cid2` (:137-144).  Here:
//
//  * the scan is replaced by a map  value -> (row, column)  over the cids held in the rows (a cid sits in at
//    most one cell unless the reference's own quirks duplicated it, in which case the entry says "ambiguous"
//    and that connection falls back to the scan), so a connection finds its rows with two shared-memory reads;
//  * within ONE limb the connections carry distinct cid1 and distinct cid2 (the greedy step uses every peak at
//    most once per limb side), so they can only interact through a row that two of them match.  One lane per
//    connection looks its rows up at the start of the limb and claims them.  A connection is SIMPLE when it
//    matches at most one row and nobody else claims that row; all simple connections are applied at once --
//    extend (:146-151) or start a row (:173-183).  The others (two matched rows: a merge or the `found == 2`
//    extend, or a shared row; in crowds a handful on the ear limbs 17-18) are COMPLEX and are walked one by one
//    in connection order with the reference's scan, after the simple extends and before the new rows are
//    appended.  Why that order is the sequential result is argued at assemble_limb below;
//  * quirks that are part of the observable behaviour are reproduced: rows hold cids as floats, the merge test
//    is `> 0` (cid 0 counts as absent) and its `+= other + 1` arithmetic (:158-160), a connection matching three
//    or more rows is dropped, limb 18 never starts a person, and peak scores are looked up by cid in the
//    part-sorted table (the sums arrive precomputed from paf_connect_kernel).
//
// All warps of the block stage the image's connections (one flat coalesced pass), warp 0 runs the 19 limbs,
// all warps prune and write the packed result record (ResultLayout): one device-to-host copy per batch.
#include "common.cuh"

namespace ekp {

// One connection as the assembly uses it (pafprocess.h:45-51 + the two score sums the reference forms at
// pafprocess.cpp:150/171 and :179-181, same operation order; formed by paf_connect_kernel).
struct __align__(16) ConnRec {
    int cid1, cid2;
    float score;    // connection score
    float s_ext;    // peak_score(cid2) + score                         (a row is extended by part2)
    float s_new;    // (peak_score(cid1) + peak_score(cid2)) + score    (a new row)
    int pad0, pad1, pad2;
};

__device__ __forceinline__ ConnRec make_rec(const Conn& cn) {
    ConnRec r;
    r.cid1 = cn.cid1; r.cid2 = cn.cid2;
    r.score = cn.score; r.s_ext = cn.s_ext; r.s_new = cn.s_new;
    r.pad0 = r.pad1 = r.pad2 = 0;
    return r;
}

constexpr int kRS = 21;  // shared-memory stride of a subset row (20 values; an odd stride spreads the lanes' rows over the banks)
enum { CLS_NOP = 0, CLS_SIMPLE1 = 1, CLS_NEW = 2, CLS_COMPLEX = 3 };
constexpr int MAP_NONE = -1, MAP_AMBIGUOUS = -2;

struct AsmState {
    float* rows;         // [max(max_humans, 32)][kRS]
    const ConnRec* sRec; // staged records (or nullptr -> build from `conns` on the fly)
    int* sMap;           // [map_n] value -> (column << 16 | row), MAP_NONE, MAP_AMBIGUOUS
    int* sOwn1;          // [max_humans] which connection of which limb claimed this row through part1 / part2:
    int* sOwn2;          //              ((limb + 1) << 16 | k); sOwn1 doubles as the kept-row list of the prune
    short* sRow;         // [2][max_part] row matched through part1 / part2 (-1: none)
    unsigned char* sCls; // [max_part]
    const int* sStart;   // [20] prefix of per-limb counts
    const Conn* conns;   // this image's [19][max_part]
    int max_part, max_humans, map_n;
    __device__ __forceinline__ ConnRec rec_at(int limb, int k) const {
        return sRec ? sRec[sStart[limb] + k] : make_rec(conns[(size_t) limb * max_part + k]);
    }
    // which row holds value v in column col?  (>= 0 row, MAP_NONE, MAP_AMBIGUOUS)
    __device__ __forceinline__ int query(int col, int v) const {
        const int e = sMap[v];
        if (e < 0) return e;
        return (e >> 16) == col ? (e & 0xffff) : MAP_NONE;
    }
    __device__ __forceinline__ void map_set(int v, int row, int col) const {
        if (v < 0 || v >= map_n) return;  // a value the merge quirk produced beyond the peak table can never be matched
        const int e = sMap[v];
        sMap[v] = e == MAP_NONE ? ((col << 16) | row) : MAP_AMBIGUOUS;
    }
    __device__ __forceinline__ void map_remove(int v, int row, int col) const {
        if (v < 0 || v >= map_n) return;
        const int e = sMap[v];
        if (e == ((col << 16) | row)) sMap[v] = MAP_NONE;
        else if (e != MAP_AMBIGUOUS) sMap[v] = MAP_AMBIGUOUS;  // never expected; the scan then decides
    }
};

// Rebuild the map from the rows (after the complex walk changed them in ways the map does not track).
__device__ void rebuild_map(const AsmState& S, int nrows) {
    const int lane = threadIdx.x;
    for (int v = lane; v < S.map_n; v += 32) S.sMap[v] = MAP_NONE;
    __syncwarp();
    for (int idx = lane; idx < nrows * EKP_NUM_PART; idx += 32) {
        const int r = idx / EKP_NUM_PART, c = idx - r * EKP_NUM_PART;
        const float f = S.rows[r * kRS + c];
        if (!(f >= 0.f) || f >= (float) S.map_n) continue;
        const int v = (int) f;
        if (atomicCAS(&S.sMap[v], MAP_NONE, (c << 16) | r) != MAP_NONE) S.sMap[v] = MAP_AMBIGUOUS;
    }
    __syncwarp();
}

// ---- one connection the reference's way: scan all rows (pafprocess.cpp:135-183), whole warp ----------------
// Returns true when the connection found no row (the caller starts a new one later, in connection order).
__device__ bool connection_by_scan(const AsmState& S, const ConnRec cn, int p1, int p2, int& nrows, bool& modified /* a merge happened */) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    float* rows = S.rows;
    const float f1 = (float) cn.cid1, f2 = (float) cn.cid2;
    int found = 0, s1 = 0, s2 = 0;
    for (int base = 0; base < nrows; base += 32) {
        const int r = base + lane;
        const bool m = r < nrows && (rows[r * kRS + p1] == f1 || rows[r * kRS + p2] == f2);
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
    auto extend_s1 = [&]() {  // rows[s1][p2] = cid2, count + 1, score + (peak(cid2) + conn)
        if (lane == 0) {
            const float old = rows[s1 * kRS + p2];
            if (old >= 0.f) S.map_remove((int) old, s1, p2);
            S.map_set(cn.cid2, s1, p2);  // ambiguous when another row holds cid2 already (the found == 2 case)
            rows[s1 * kRS + p2] = f2;
            rows[s1 * kRS + 19] = __fadd_rn(rows[s1 * kRS + 19], 1.0f);
            rows[s1 * kRS + 18] = __fadd_rn(rows[s1 * kRS + 18], cn.s_ext);
        }
    };
    bool starts = false;
    if (found == 1) {
        if (rows[s1 * kRS + p2] != f2) extend_s1();
    } else if (found == 2) {
        const bool both = lane < 18 && rows[s1 * kRS + lane] > 0.f && rows[s2 * kRS + lane] > 0.f;
        const bool membership = __any_sync(FULL, both);
        if (!membership) {
            if (lane < 18) rows[s1 * kRS + lane] = __fadd_rn(rows[s1 * kRS + lane], __fadd_rn(rows[s2 * kRS + lane], 1.0f));
            if (lane == 19) rows[s1 * kRS + 19] = __fadd_rn(rows[s1 * kRS + 19], rows[s2 * kRS + 19]);
            if (lane == 18) {
                const float v = __fadd_rn(rows[s1 * kRS + 18], rows[s2 * kRS + 18]);
                rows[s1 * kRS + 18] = __fadd_rn(v, cn.score);
            }
            __syncwarp();
            if (lane < 20)  // erase row s2: every lane shifts its own column
                for (int r = s2; r < nrows - 1; r++) rows[r * kRS + lane] = rows[(r + 1) * kRS + lane];
            nrows--;
            modified = true;  // rows moved: the caller rebuilds the map
        } else {
            extend_s1();
        }
    } else if (found == 0) {
        starts = true;
    }
    __syncwarp();
    return starts;
}

#ifdef EKP_ASM_PROFILE  // tools/ only: time per phase summed over images (ns)
__device__ unsigned long long g_asm_prof[8];
__device__ __forceinline__ unsigned long long asm_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define APROF(k) do { if (threadIdx.x == 0) { const unsigned long long _t = asm_now(); atomicAdd(&g_asm_prof[k], _t - prof_t); prof_t = _t; } } while (0)
#define APROF_DECL unsigned long long prof_t = asm_now()
#define ACOUNT(k, v) do { if (threadIdx.x == 0) atomicAdd(&g_asm_prof[k], (unsigned long long) (v)); } while (0)
extern "C" int ekp_debug_asm_profile(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, g_asm_prof, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_asm_prof, z, sizeof(z)); }
    return 0;
}
#else
#define APROF(k) do { } while (0)
#define APROF_DECL do { } while (0)
#define ACOUNT(k, v) do { } while (0)
#endif

// ---- one limb -----------------------------------------------------------------------------------------------
// Let the limb's connections be k = 0..nc-1 (acceptance order) with distinct cid1 and distinct cid2, T_k the set
// of rows k matches at the START of the limb, and C the union of the T_k of all COMPLEX connections (|T_k| = 2,
// or a row of T_k is also in some other T_j, or the map could not answer).  Claims:
//  (1) A SIMPLE connection (|T_k| <= 1, its row claimed by nobody else) takes, in the sequential walk, the same
//      branch on the same row as at the start of the limb, and what it does is invisible to every other
//      connection: it only writes its own row r (not in C, not in any other T_j): r[p2] = cid2_k (no other
//      connection carries that cid2), r[18], r[19]; the value it overwrites in r[p2] would put r into the T_j of
//      the connection carrying it, which contradicts "claimed by nobody else".
//  (2) Whatever the complex connections do stays inside C: an extend writes a row of its current match set, a
//      merge moves values between two rows of it, and the current match set of a complex connection only ever
//      contains rows that held cid1 / cid2 at the start (in C) or rows those values were moved to (in C).
//  (3) A row started by this limb holds cid1_k / cid2_k, which no other connection of the limb carries, so it is
//      never matched before the limb ends; the reference appends such rows in connection order and erases merged
//      rows in place, so the final row order is: surviving old rows in their order, then the new rows in
//      connection order.
// Hence: simple extends (parallel), then the complex connections one by one with the reference's scan over the
// old rows, then all new rows appended in connection order -- the reference's result, bit for bit.
__device__ void assemble_limb(const AsmState& S, int limb, int& nrows_io, bool& ovf) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
    const int nc = S.sStart[limb + 1] - S.sStart[limb];
    float* rows = S.rows;
    int nrows = nrows_io;
    APROF_DECL;
    // A: look the rows up, claim them.  The cid1 of a limb's connections are distinct and so are the cid2, hence two
    // connections can only meet in a row that one of them matches through part1 and the other through part2: every
    // connection notes itself in sOwn1[row matched through part1] / sOwn2[row matched through part2] (tagged with the
    // limb, so nothing has to be cleared) and afterwards looks for a foreign note in the OTHER table.
    const int tag0 = (limb + 1) << 16;
    bool any_complex = false;
    for (int k = lane; k < nc; k += 32) {
        const ConnRec cn = S.rec_at(limb, k);
        const int q1 = S.query(p1, cn.cid1), q2 = S.query(p2, cn.cid2);
        int cls;
        if (q1 == MAP_AMBIGUOUS || q2 == MAP_AMBIGUOUS) {  // duplicated cid: claim what the reference's scan would match
            const float f1 = (float) cn.cid1, f2 = (float) cn.cid2;
            for (int r = 0; r < nrows; r++)
                if (rows[r * kRS + p1] == f1 || rows[r * kRS + p2] == f2) { S.sOwn1[r] = tag0 | 0xffff; S.sOwn2[r] = tag0 | 0xffff; }
            cls = CLS_COMPLEX;
        } else {
            if (q1 >= 0) S.sOwn1[q1] = tag0 | k;
            if (q2 >= 0) S.sOwn2[q2] = tag0 | k;
            if (q1 >= 0 && q2 >= 0 && q1 != q2) cls = CLS_COMPLEX;  // found == 2
            else if (q1 >= 0 || q2 >= 0) {
                // matched through part2 only, or both cids already in this row: the reference changes nothing (:147)
                cls = (q1 >= 0 && q2 < 0) ? CLS_SIMPLE1 : CLS_NOP;
            } else {
                cls = limb < 18 ? CLS_NEW : CLS_NOP;  // found == 0 (:173): limb 18 never starts a person
            }
        }
        S.sRow[k] = (short) (q1 >= 0 ? q1 : -1);
        S.sRow[S.max_part + k] = (short) (q2 >= 0 ? q2 : -1);
        S.sCls[k] = (unsigned char) cls;
    }
    __syncwarp();
    // B: a row claimed by two connections makes both complex; the simple extends are applied
    for (int k = lane; k < nc; k += 32) {
        int cls = S.sCls[k];
        const int q1 = S.sRow[k], q2 = S.sRow[S.max_part + k];
        if (cls != CLS_COMPLEX) {
            const int o2 = q1 >= 0 ? S.sOwn2[q1] : 0, o1 = q2 >= 0 ? S.sOwn1[q2] : 0;
            const bool foreign = ((o2 >> 16) == limb + 1 && (o2 & 0xffff) != k) || ((o1 >> 16) == limb + 1 && (o1 & 0xffff) != k);
            if (foreign) { cls = CLS_COMPLEX; S.sCls[k] = CLS_COMPLEX; }
        }
        if (cls == CLS_SIMPLE1) {  // found == 1 through part1, row[p2] != cid2: pafprocess.cpp:146-151
            const ConnRec cn = S.rec_at(limb, k);
            const int row = q1;
            const float old = rows[row * kRS + p2];
            rows[row * kRS + p2] = (float) cn.cid2;
            rows[row * kRS + 19] = __fadd_rn(rows[row * kRS + 19], 1.0f);
            rows[row * kRS + 18] = __fadd_rn(rows[row * kRS + 18], cn.s_ext);
            if (old >= 0.f) S.map_remove((int) old, row, p2);
            S.map_set(cn.cid2, row, p2);
        }
        any_complex |= cls == CLS_COMPLEX;
    }
    __syncwarp();
    APROF(3);  // phases A + B
    // C: the complex connections, in connection order, the reference's way
    if (__any_sync(FULL, any_complex)) {
        bool modified = false;
        for (int k0 = 0; k0 < nc; k0 += 32) {
            const int k = k0 + lane;
            unsigned mask = __ballot_sync(FULL, k < nc && S.sCls[k] == CLS_COMPLEX);
            while (mask) {
                const int kk = k0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const ConnRec cn = S.rec_at(limb, kk);
                ACOUNT(6, 1);
                const bool starts = connection_by_scan(S, cn, p1, p2, nrows, modified);
                if (lane == 0) S.sCls[kk] = (unsigned char) (starts && limb < 18 ? CLS_NEW : CLS_NOP);
                __syncwarp();
            }
        }
        if (modified) { rebuild_map(S, nrows); ACOUNT(7, 1); }
    }
    APROF(4);  // phase C
    // D: new rows, in connection order (:173-183)
    int nnew = 0;
    for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        const bool starts = k < nc && S.sCls[k] == CLS_NEW;
        const unsigned mask = __ballot_sync(FULL, starts);
        if (starts) {
            const int r = nrows + nnew + __popc(mask & ((1u << lane) - 1u));
            if (r < S.max_humans) {
                const ConnRec cn = S.rec_at(limb, k);
#pragma unroll
                for (int q = 0; q < 18; q++) rows[r * kRS + q] = -1.0f;
                rows[r * kRS + p1] = (float) cn.cid1;
                rows[r * kRS + p2] = (float) cn.cid2;
                rows[r * kRS + 18] = cn.s_new;
                rows[r * kRS + 19] = 2.0f;
                S.map_set(cn.cid1, r, p1);
                S.map_set(cn.cid2, r, p2);
            }
        }
        nnew += __popc(mask);
    }
    if (nrows + nnew > S.max_humans) { ovf = true; nnew = S.max_humans - nrows; }
    nrows_io = nrows + nnew;
    __syncwarp();
    APROF(5);  // phase D
}

// The same limb with up to K connections per lane (k = lane + 32 q) held in registers: no round trips through shared
// memory between the phases, and the K lookups of a lane are independent (the usual case: nc <= 64).
template <int K>
__device__ __forceinline__ void assemble_limb_regs(const AsmState& S, int limb, int& nrows_io, bool& ovf) {
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int p1 = kPairs[limb][0], p2 = kPairs[limb][1];
    const int nc = S.sStart[limb + 1] - S.sStart[limb];
    float* rows = S.rows;
    int nrows = nrows_io;
    APROF_DECL;
    const int tag0 = (limb + 1) << 16;
    ConnRec cn[K];
    int q1[K], q2[K], cls[K];
    bool act[K];
#pragma unroll
    for (int q = 0; q < K; q++) {
        act[q] = lane + 32 * q < nc;
        if (act[q]) cn[q] = S.rec_at(limb, lane + 32 * q);
    }
    // A: look the rows up, claim them (see assemble_limb)
#pragma unroll
    for (int q = 0; q < K; q++) {
        q1[q] = q2[q] = MAP_NONE;
        cls[q] = CLS_NOP;
        if (!act[q]) continue;
        q1[q] = S.query(p1, cn[q].cid1);
        q2[q] = S.query(p2, cn[q].cid2);
    }
#pragma unroll
    for (int q = 0; q < K; q++) {
        if (!act[q]) continue;
        const int k = lane + 32 * q;
        if (q1[q] == MAP_AMBIGUOUS || q2[q] == MAP_AMBIGUOUS) {
            const float f1 = (float) cn[q].cid1, f2 = (float) cn[q].cid2;
            for (int r = 0; r < nrows; r++)
                if (rows[r * kRS + p1] == f1 || rows[r * kRS + p2] == f2) { S.sOwn1[r] = tag0 | 0xffff; S.sOwn2[r] = tag0 | 0xffff; }
            cls[q] = CLS_COMPLEX;
            q1[q] = q2[q] = MAP_NONE;
        } else {
            if (q1[q] >= 0) S.sOwn1[q1[q]] = tag0 | k;
            if (q2[q] >= 0) S.sOwn2[q2[q]] = tag0 | k;
            if (q1[q] >= 0 && q2[q] >= 0 && q1[q] != q2[q]) cls[q] = CLS_COMPLEX;           // found == 2
            else if (q1[q] >= 0 || q2[q] >= 0) cls[q] = (q1[q] >= 0 && q2[q] < 0) ? CLS_SIMPLE1 : CLS_NOP;
            else cls[q] = limb < 18 ? CLS_NEW : CLS_NOP;                                    // found == 0
        }
    }
    __syncwarp();
    // B: foreign claims make a connection complex; the simple extends are applied
    bool any_complex = false;
#pragma unroll
    for (int q = 0; q < K; q++) {
        if (!act[q]) continue;
        const int k = lane + 32 * q;
        if (cls[q] != CLS_COMPLEX) {
            const int o2 = q1[q] >= 0 ? S.sOwn2[q1[q]] : 0, o1 = q2[q] >= 0 ? S.sOwn1[q2[q]] : 0;
            if (((o2 >> 16) == limb + 1 && (o2 & 0xffff) != k) || ((o1 >> 16) == limb + 1 && (o1 & 0xffff) != k)) cls[q] = CLS_COMPLEX;
        }
        if (cls[q] == CLS_SIMPLE1) {  // found == 1 through part1, row[p2] != cid2: pafprocess.cpp:146-151
            const int row = q1[q];
            const float old = rows[row * kRS + p2];
            rows[row * kRS + p2] = (float) cn[q].cid2;
            rows[row * kRS + 19] = __fadd_rn(rows[row * kRS + 19], 1.0f);
            rows[row * kRS + 18] = __fadd_rn(rows[row * kRS + 18], cn[q].s_ext);
            if (old >= 0.f) S.map_remove((int) old, row, p2);
            S.map_set(cn[q].cid2, row, p2);
        }
        any_complex |= cls[q] == CLS_COMPLEX;
    }
    __syncwarp();
    APROF(3);  // phases A + B
    // C: the complex connections, in connection order, the reference's way
    if (__any_sync(FULL, any_complex)) {
        bool modified = false;
#pragma unroll
        for (int q = 0; q < K; q++) {
            unsigned mask = __ballot_sync(FULL, act[q] && cls[q] == CLS_COMPLEX);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                ConnRec c;
                c.cid1 = __shfl_sync(FULL, cn[q].cid1, src); c.cid2 = __shfl_sync(FULL, cn[q].cid2, src);
                c.score = __shfl_sync(FULL, cn[q].score, src); c.s_ext = __shfl_sync(FULL, cn[q].s_ext, src);
                c.s_new = __shfl_sync(FULL, cn[q].s_new, src);
                ACOUNT(6, 1);
                const bool starts = connection_by_scan(S, c, p1, p2, nrows, modified);
                if (lane == src) cls[q] = starts && limb < 18 ? CLS_NEW : CLS_NOP;
            }
        }
        if (modified) { rebuild_map(S, nrows); ACOUNT(7, 1); }
    }
    APROF(4);  // phase C
    // D: new rows, in connection order (:173-183)
    int nnew = 0;
#pragma unroll
    for (int q = 0; q < K; q++) {
        const bool starts = act[q] && cls[q] == CLS_NEW;
        const unsigned mask = __ballot_sync(FULL, starts);
        if (starts) {
            const int r = nrows + nnew + __popc(mask & ((1u << lane) - 1u));
            if (r < S.max_humans) {
#pragma unroll
                for (int c = 0; c < 18; c++) rows[r * kRS + c] = -1.0f;
                rows[r * kRS + p1] = (float) cn[q].cid1;
                rows[r * kRS + p2] = (float) cn[q].cid2;
                rows[r * kRS + 18] = cn[q].s_new;
                rows[r * kRS + 19] = 2.0f;
                S.map_set(cn[q].cid1, r, p1);
                S.map_set(cn[q].cid2, r, p2);
            }
        }
        nnew += __popc(mask);
    }
    if (nrows + nnew > S.max_humans) { ovf = true; nnew = S.max_humans - nrows; }
    nrows_io = nrows + nnew;
    __syncwarp();
    APROF(5);  // phase D
}

size_t assemble_smem_bytes(int max_humans, int max_peaks, int max_part, int conn_cap) {
    const size_t nrow = (size_t) (max_humans < 32 ? 32 : max_humans);
    size_t b = sizeof(float) * kRS * nrow;                // rows
    b = (b + 15) & ~(size_t) 15;
    b += sizeof(ConnRec) * (size_t) conn_cap;             // sRec
    b += sizeof(int) * (size_t) max_peaks;                // sMap
    b += sizeof(int) * 2 * nrow;                          // sOwn1 / sKept, sOwn2
    b += sizeof(short) * 2 * (size_t) max_part;           // sRow
    b += (size_t) max_part;                               // sCls
    return (b + 15) & ~(size_t) 15;
}
int assemble_conn_cap(int max_humans, int max_part) {
    const long long want = 24ll * max_humans, all = 19ll * max_part;
    long long c = want < all ? want : all;
    if (c > 1536) c = 1536;
    if (c < 64) c = 64;
    return (int) c;
}

// The whole block works on image `img`; `smem_raw` must provide assemble_smem_bytes(...) bytes, 16-byte aligned.
__device__ void assemble_image(const AsmParams& P, int img, unsigned char* smem_raw) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t nrow_cap = (size_t) (P.max_humans < 32 ? 32 : P.max_humans);
    float* rows = reinterpret_cast<float*>(smem_raw);
    ConnRec* sRec = reinterpret_cast<ConnRec*>(smem_raw + ((sizeof(float) * kRS * nrow_cap + 15) & ~(size_t) 15));
    int* sMap = reinterpret_cast<int*>(sRec + P.conn_cap);
    int* sOwn1 = sMap + P.max_peaks;
    int* sOwn2 = sOwn1 + nrow_cap;
    short* sRow = reinterpret_cast<short*>(sOwn2 + nrow_cap);
    unsigned char* sCls = reinterpret_cast<unsigned char*>(sRow + 2 * P.max_part);
    __shared__ int sStart[EKP_NUM_LIMB + 1];
    __shared__ int sNrows, sOvf;
    APROF_DECL;
    const ekp_peak* L = P.line + (size_t) img * P.max_peaks;
    const Conn* Cimg = P.conns + (size_t) img * EKP_NUM_LIMB * P.max_part;
    const int npk = P.n_peaks[img];
    const int map_n = min(P.part_off[(size_t) img * 20 + EKP_NUM_PART + 1], P.max_peaks);  // every id is below the raw peak count

    // ---- stage one record per connection (all warps), initialise the map --------------------------------
    if (tid < 32) {
        int cnt = 0;
        if (tid < EKP_NUM_LIMB) cnt = min(P.n_conns[(size_t) img * EKP_NUM_LIMB + tid], P.max_part);
        int incl = cnt;  // inclusive prefix over the 19 limbs
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        if (tid < EKP_NUM_LIMB) sStart[tid] = incl - cnt;
        if (tid == EKP_NUM_LIMB - 1) sStart[EKP_NUM_LIMB] = incl;
    }
    for (int v = tid; v < map_n; v += nthr) sMap[v] = MAP_NONE;
    for (int r = tid; r < 2 * (int) nrow_cap; r += nthr) sOwn1[r] = 0;  // both claim tables
    __syncthreads();
    const int total_conns = sStart[EKP_NUM_LIMB];
    const bool staged = total_conns <= P.conn_cap;
    if (staged) {
        // one flat pass so that all loads are in flight together (a per-limb loop would pay one
        // global-memory round trip per limb)
        int start[EKP_NUM_LIMB];  // per-limb offsets in registers: the limb of a flat index costs no memory access
#pragma unroll
        for (int l = 0; l < EKP_NUM_LIMB; l++) start[l] = sStart[l];
        for (int idx = tid; idx < total_conns; idx += nthr) {
            int limb = 0, base = 0;
#pragma unroll
            for (int l = 1; l < EKP_NUM_LIMB; l++)
                if (idx >= start[l]) { limb = l; base = start[l]; }
            sRec[idx] = make_rec(Cimg[(size_t) limb * P.max_part + (idx - base)]);
        }
    }
    __syncthreads();
    APROF(0);  // staging

    // ---- assembly, limb by limb (pafprocess.cpp:130-185): warp 0 ---------------------------------------
    if (tid < 32) {
        AsmState S;
        S.rows = rows; S.sRec = staged ? sRec : nullptr; S.sMap = sMap; S.sOwn1 = sOwn1; S.sOwn2 = sOwn2; S.sRow = sRow; S.sCls = sCls;
        S.sStart = sStart; S.conns = Cimg; S.max_part = P.max_part; S.max_humans = P.max_humans; S.map_n = map_n;
        int nrows = 0;
        bool ovf = false;
        for (int limb = 0; limb < EKP_NUM_LIMB; limb++) {
            const int nc = sStart[limb + 1] - sStart[limb];
            if (nc == 0) continue;
            if (nc <= 32) assemble_limb_regs<1>(S, limb, nrows, ovf);
            else if (nc <= 64) assemble_limb_regs<2>(S, limb, nrows, ovf);
            else assemble_limb(S, limb, nrows, ovf);
        }
        if (tid == 0) { sNrows = nrows; sOvf = ovf ? 1 : 0; }
    }
    __syncthreads();
    APROF(1);  // limbs

    // ---- prune (pafprocess.cpp:187-191: a reverse erase loop == an order-preserving filter) and
    //      write the image's result record, all threads busy ---------------------------------------
    const int nrows = sNrows;
    int* sKept = sOwn1;
    unsigned char* rec = P.records + (size_t) img * P.lay.stride;
    float* so = reinterpret_cast<float*>(rec + P.lay.off_subset);
    ekp_peak* hp = reinterpret_cast<ekp_peak*>(rec + P.lay.off_hparts);
    float* hs = reinterpret_cast<float*>(rec + P.lay.off_hscore);
    __shared__ int sKeptN;
    if (tid < 32) {
        int kept = 0;
        for (int base = 0; base < nrows; base += 32) {
            const int r = base + tid;
            bool keep = false;
            if (r < nrows) {
                const float c = rows[r * kRS + 19], sc = rows[r * kRS + 18];
                keep = !(c < 4.0f || __fdiv_rn(sc, c) < 0.3f);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) sKept[kept + __popc(mask & ((1u << tid) - 1u))] = r;
            kept += __popc(mask);
        }
        if (tid == 0) sKeptN = kept;
    }
    __syncthreads();
    const int kept = sKeptN;
    for (int idx = tid; idx < kept * 20; idx += nthr) {
        const int k = idx / 20, q = idx - k * 20;
        so[idx] = rows[sKept[k] * kRS + q];
    }
    for (int idx = tid; idx < kept * EKP_NUM_PART; idx += nthr) {
        const int k = idx / EKP_NUM_PART, q = idx - k * EKP_NUM_PART;
        const int cid = (int) rows[sKept[k] * kRS + q];  // get_part_cid: float -> int
        ekp_peak o;
        if (cid >= 0) { const ekp_peak pk = L[cid]; o.x = pk.x; o.y = pk.y; o.score = pk.score; o.id = cid; }
        else { o.x = 0; o.y = 0; o.score = 0.f; o.id = -1; }
        hp[idx] = o;
    }
    for (int k = tid; k < kept; k += nthr) {
        const int r = sKept[k];
        hs[k] = __fdiv_rn(rows[r * kRS + 18], rows[r * kRS + 19]);  // get_score
    }
    if (tid == 0) {
        int4 head;
        head.x = kept;
        head.y = npk;
        head.z = (int) (P.overflow[img] | (sOvf ? EKP_OVF_HUMANS : 0u));
        head.w = 0;
        *reinterpret_cast<int4*>(rec) = head;
    }
    APROF(2);  // prune + record
}

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
