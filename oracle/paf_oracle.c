/* oracle/paf_oracle.c -- TEST INFRASTRUCTURE, never part of the product path.
 *
 * Plain-C restatement of the reference's PAF post-processing native op
 * (/root/reference/lib/pafprocess/pafprocess.cpp:22-246, constants and tables
 * from pafprocess.h:6-24).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's .so.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors (SURVEY.md
 * section 4), so this restatement is pinned against the reference itself:
 * oracle/_ref/libpaf_ref.so (the unmodified reference C++ compiled from
 * /root/reference by oracle/Makefile) on seeded scenes, and against the
 * committed fixtures tests/golden/ (.npz) that were generated from that compiled
 * reference (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off (no -ffast-math, no -march=native): the
 * reference is built by distutils for baseline x86-64, so no FMA contraction.
 *
 * The candidate sort is libstdc++'s std::sort (GCC 13, bits/stl_algo.h
 * __sort/__introsort_loop/__final_insertion_sort, bits/stl_heap.h for the
 * depth-limit fallback).  It is a third-party ORDERING dependency that is not
 * under /root/reference: its tie permutation is observable in the reference's
 * results (pafprocess.cpp:97), so the published algorithm is restated here
 * (okp_sort_*), threshold 16, depth limit 2*floor(log2 n).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OKP_NUM_PART 18
#define OKP_NUM_LIMB 19
#define OKP_STEP_PAF 10

static const float OKP_THRESH_VECTOR_SCORE = 0.05; /* pafprocess.h:7, const float = (float)0.05 */
static const int OKP_THRESH_VECTOR_CNT1 = 6;        /* pafprocess.h:8  */
static const int OKP_THRESH_PART_CNT = 4;           /* pafprocess.h:9  */
static const float OKP_THRESH_HUMAN_SCORE = 0.3;    /* pafprocess.h:10 */

/* pafprocess.h:16-19 */
static const int OKP_PAIRS_NET[OKP_NUM_LIMB][2] = {
    {12, 13}, {20, 21}, {14, 15}, {16, 17}, {22, 23}, {24, 25}, {0, 1},   {2, 3},   {4, 5},   {6, 7},
    {8, 9},   {10, 11}, {28, 29}, {30, 31}, {34, 35}, {32, 33}, {36, 37}, {18, 19}, {26, 27}};
/* pafprocess.h:21-24 */
static const int OKP_PAIRS[OKP_NUM_LIMB][2] = {
    {1, 2},   {1, 5},   {2, 3},  {3, 4},   {5, 6},   {6, 7},  {1, 8},   {8, 9},  {9, 10}, {1, 11},
    {11, 12}, {12, 13}, {1, 0},  {0, 14},  {14, 16}, {0, 15}, {15, 17}, {2, 16}, {5, 17}};

typedef struct { int x, y; float score; int id; } okp_peak;                /* pafprocess.h:26-31 */
typedef struct { int idx1, idx2; float score, etc; } okp_cand;             /* pafprocess.h:38-43 */
typedef struct { int cid1, cid2; float score; int peak_id1, peak_id2; } okp_conn; /* pafprocess.h:45-51 */

/* ---- growable arrays ------------------------------------------------------ */
#define OKP_VEC(T, name)                                                              \
    typedef struct { T *v; int n, cap; } name;                                        \
    static void name##_push(name *a, T e) {                                           \
        if (a->n == a->cap) {                                                         \
            a->cap = a->cap ? a->cap * 2 : 16;                                        \
            a->v = (T *) realloc(a->v, sizeof(T) * (size_t) a->cap);                  \
        }                                                                             \
        a->v[a->n++] = e;                                                             \
    }                                                                                 \
    static void name##_clear(name *a) { a->n = 0; }
OKP_VEC(okp_peak, peakvec)
OKP_VEC(okp_cand, candvec)
OKP_VEC(okp_conn, connvec)
typedef struct { float r[20]; } okp_row;
OKP_VEC(okp_row, rowvec)

/* process-global result state, as in the reference (pafprocess.cpp:12-13) */
static rowvec g_subset;
static peakvec g_peaks_line;
static connvec g_conns[OKP_NUM_LIMB];   /* kept for finer-grained diffs */
static candvec g_cands[OKP_NUM_LIMB];   /* sorted candidates, ditto */

/* ---- libstdc++ std::sort restatement -------------------------------------- */
static int okp_comp(const okp_cand *a, const okp_cand *b) { return a->score > b->score; } /* pafprocess.cpp:244-246 */
static void okp_swap(okp_cand *a, okp_cand *b) { okp_cand t = *a; *a = *b; *b = t; }

static void okp_unguarded_linear_insert(okp_cand *last) {
    okp_cand val = *last;
    okp_cand *next = last - 1;
    while (okp_comp(&val, next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void okp_insertion_sort(okp_cand *first, okp_cand *last) {
    if (first == last) return;
    for (okp_cand *i = first + 1; i != last; ++i) {
        if (okp_comp(i, first)) {
            okp_cand val = *i;
            memmove(first + 1, first, (size_t) (i - first) * sizeof(okp_cand));
            *first = val;
        } else {
            okp_unguarded_linear_insert(i);
        }
    }
}
static void okp_unguarded_insertion_sort(okp_cand *first, okp_cand *last) {
    for (okp_cand *i = first; i != last; ++i) okp_unguarded_linear_insert(i);
}
static void okp_final_insertion_sort(okp_cand *first, okp_cand *last) {
    if (last - first > 16) {
        okp_insertion_sort(first, first + 16);
        okp_unguarded_insertion_sort(first + 16, last);
    } else {
        okp_insertion_sort(first, last);
    }
}
static void okp_move_median_to_first(okp_cand *result, okp_cand *a, okp_cand *b, okp_cand *c) {
    if (okp_comp(a, b)) {
        if (okp_comp(b, c)) okp_swap(result, b);
        else if (okp_comp(a, c)) okp_swap(result, c);
        else okp_swap(result, a);
    } else if (okp_comp(a, c)) okp_swap(result, a);
    else if (okp_comp(b, c)) okp_swap(result, c);
    else okp_swap(result, b);
}
static okp_cand *okp_unguarded_partition(okp_cand *first, okp_cand *last, okp_cand *pivot) {
    for (;;) {
        while (okp_comp(first, pivot)) ++first;
        --last;
        while (okp_comp(pivot, last)) --last;
        if (!(first < last)) return first;
        okp_swap(first, last);
        ++first;
    }
}
/* heap fallback (bits/stl_heap.h) */
static void okp_push_heap(okp_cand *first, long hole, long top, okp_cand value) {
    long parent = (hole - 1) / 2;
    while (hole > top && okp_comp(first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void okp_adjust_heap(okp_cand *first, long hole, long len, okp_cand value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (okp_comp(first + child, first + (child - 1))) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    okp_push_heap(first, hole, top, value);
}
static void okp_pop_heap(okp_cand *first, okp_cand *last, okp_cand *result) {
    okp_cand value = *result;
    *result = *first;
    okp_adjust_heap(first, 0, last - first, value);
}
static void okp_make_heap(okp_cand *first, okp_cand *last) {
    long len = last - first;
    if (len < 2) return;
    long parent = (len - 2) / 2;
    for (;;) {
        okp_cand value = first[parent];
        okp_adjust_heap(first, parent, len, value);
        if (parent == 0) return;
        parent--;
    }
}
static void okp_partial_sort_all(okp_cand *first, okp_cand *last) {
    /* __partial_sort(first, last, last): __heap_select then __sort_heap */
    okp_make_heap(first, last); /* the i in [middle,last) loop is empty: middle == last */
    while (last - first > 1) { --last; okp_pop_heap(first, last, last); }
}
static int g_heapsort_hits = 0; /* how often the depth-limit fallback ran (test visibility) */
static void okp_introsort_loop(okp_cand *first, okp_cand *last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) { g_heapsort_hits++; okp_partial_sort_all(first, last); return; }
        --depth_limit;
        okp_cand *mid = first + (last - first) / 2;
        okp_move_median_to_first(first, first + 1, mid, last - 1);
        okp_cand *cut = okp_unguarded_partition(first + 1, last, first);
        okp_introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}
static long okp_lg(long n) { long k = 0; while (n > 1) { n >>= 1; k++; } return k; }
static void okp_sort(okp_cand *first, okp_cand *last) {
    if (first != last) {
        okp_introsort_loop(first, last, okp_lg(last - first) * 2);
        okp_final_insertion_sort(first, last);
    }
}

/* exported for the sort unit tests: sorts n (score, tag) pairs, tag rides in idx1 */
void okp_sort_scores(int n, float *score, int *tag) {
    okp_cand *c = (okp_cand *) malloc(sizeof(okp_cand) * (size_t) (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { c[i].idx1 = tag[i]; c[i].idx2 = 0; c[i].score = score[i]; c[i].etc = 0; }
    okp_sort(c, c + n);
    for (int i = 0; i < n; i++) { score[i] = c[i].score; tag[i] = c[i].idx1; }
    free(c);
}
int okp_heapsort_hits(void) { return g_heapsort_hits; }

/* ---- process_paf ----------------------------------------------------------- */
static int okp_roundpaf(float v) { return (int) ((double) v + 0.5); } /* pafprocess.cpp:240-242 */

/* Returns 0 like the reference (pafprocess.cpp:193); -1 if a part id is out of
 * range (the reference has undefined behaviour there; the oracle refuses). */
int okp_process_paf(int p1, int p2, int p3, const float *peaks, int h1, int h2, int h3, const float *heatmap,
                    int f1, int f2, int f3, const float *pafmap) {
    (void) h2; (void) h3; (void) heatmap; (void) f1; /* heat is used only through h1 (pafprocess.cpp:83) */
    static peakvec peak_infos[OKP_NUM_PART];
    for (int k = 0; k < OKP_NUM_PART; k++) peakvec_clear(&peak_infos[k]);

    /* ingest, pafprocess.cpp:24-36 */
    int peak_cnt = 0;
    for (int img = 0; img < p1; img++) {
        for (int k = 0; k < p2; k++) {
            const float *row = peaks + (size_t) p3 * ((size_t) k + (size_t) p2 * (size_t) img);
            okp_peak info;
            info.id = peak_cnt++;
            info.x = (int) row[0];
            info.y = (int) row[1];
            info.score = row[2];
            int part_id = (int) row[4];
            if (part_id < 0 || part_id >= OKP_NUM_PART) return -1;
            peakvec_push(&peak_infos[part_id], info);
        }
    }
    /* flatten part-major, pafprocess.cpp:38-43 */
    peakvec_clear(&g_peaks_line);
    for (int part = 0; part < OKP_NUM_PART; part++)
        for (int i = 0; i < peak_infos[part].n; i++) peakvec_push(&g_peaks_line, peak_infos[part].v[i]);

    /* connections per limb, pafprocess.cpp:46-125 */
    for (int pair_id = 0; pair_id < OKP_NUM_LIMB; pair_id++) {
        candvec *cands = &g_cands[pair_id];
        connvec *conns = &g_conns[pair_id];
        candvec_clear(cands);
        connvec_clear(conns);
        peakvec *A = &peak_infos[OKP_PAIRS[pair_id][0]];
        peakvec *B = &peak_infos[OKP_PAIRS[pair_id][1]];
        if (A->n == 0 || B->n == 0) continue;
        const int ch1 = OKP_PAIRS_NET[pair_id][0], ch2 = OKP_PAIRS_NET[pair_id][1];

        for (int ia = 0; ia < A->n; ia++) {
            const okp_peak *a = &A->v[ia];
            for (int ib = 0; ib < B->n; ib++) {
                const okp_peak *b = &B->v[ib];
                float vx = (float) (b->x - a->x);
                float vy = (float) (b->y - a->y);
                float norm = sqrtf(vx * vx + vy * vy); /* (float)sqrt(float): identical rounding */
                if ((double) norm < 1e-12) continue;
                vx = vx / norm;
                vy = vy / norm;

                /* get_paf_vectors, pafprocess.cpp:220-238 */
                const float step_x = (float) (b->x - a->x) / (float) OKP_STEP_PAF;
                const float step_y = (float) (b->y - a->y) / (float) OKP_STEP_PAF;
                float scores = 0.0f;
                int criterion1 = 0;
                for (int i = 0; i < OKP_STEP_PAF; i++) {
                    int lx = okp_roundpaf((float) a->x + (float) i * step_x);
                    int ly = okp_roundpaf((float) a->y + (float) i * step_y);
                    size_t base = (size_t) f3 * ((size_t) lx + (size_t) f2 * (size_t) ly);
                    float px = pafmap[base + (size_t) ch1];
                    float py = pafmap[base + (size_t) ch2];
                    float score = vx * px + vy * py;
                    scores += score;
                    if (score > OKP_THRESH_VECTOR_SCORE) criterion1 += 1;
                }
                double penalty = 0.5 * (double) h1 / (double) norm - 1.0;
                double mn = (penalty < 0.0) ? penalty : 0.0; /* std::min(0.0, penalty) */
                float criterion2 = (float) ((double) (scores / (float) OKP_STEP_PAF) + mn);

                if (criterion1 > OKP_THRESH_VECTOR_CNT1 && criterion2 > 0) {
                    okp_cand c;
                    c.idx1 = ia;
                    c.idx2 = ib;
                    c.score = criterion2;
                    c.etc = criterion2 + a->score + b->score;
                    candvec_push(cands, c);
                }
            }
        }

        okp_sort(cands->v, cands->v + cands->n); /* pafprocess.cpp:97 */
        for (int c_id = 0; c_id < cands->n; c_id++) { /* greedy, pafprocess.cpp:98-124 */
            const okp_cand *c = &cands->v[c_id];
            int assigned = 0;
            for (int k = 0; k < conns->n; k++) {
                if (conns->v[k].peak_id1 == c->idx1 || conns->v[k].peak_id2 == c->idx2) { assigned = 1; break; }
            }
            if (assigned) continue;
            okp_conn conn;
            conn.peak_id1 = c->idx1;
            conn.peak_id2 = c->idx2;
            conn.score = c->score;
            conn.cid1 = A->v[c->idx1].id;
            conn.cid2 = B->v[c->idx2].id;
            connvec_push(conns, conn);
        }
    }

    /* subset assembly, pafprocess.cpp:127-185 */
    rowvec_clear(&g_subset);
    for (int pair_id = 0; pair_id < OKP_NUM_LIMB; pair_id++) {
        connvec *conns = &g_conns[pair_id];
        const int part1 = OKP_PAIRS[pair_id][0], part2 = OKP_PAIRS[pair_id][1];
        for (int k = 0; k < conns->n; k++) {
            const okp_conn *cn = &conns->v[k];
            int found = 0, s1 = 0, s2 = 0;
            for (int s = 0; s < g_subset.n; s++) {
                if (g_subset.v[s].r[part1] == (float) cn->cid1 || g_subset.v[s].r[part2] == (float) cn->cid2) {
                    if (found == 0) s1 = s;
                    if (found == 1) s2 = s;
                    found += 1;
                }
            }
            if (found == 1) {
                float *r = g_subset.v[s1].r;
                if (r[part2] != (float) cn->cid2) {
                    r[part2] = (float) cn->cid2;
                    r[19] += 1;
                    r[18] += g_peaks_line.v[cn->cid2].score + cn->score;
                }
            } else if (found == 2) {
                float *r1 = g_subset.v[s1].r, *r2 = g_subset.v[s2].r;
                int membership = 0;
                for (int q = 0; q < 18; q++)
                    if (r1[q] > 0 && r2[q] > 0) membership = 2;
                if (membership == 0) {
                    for (int q = 0; q < 18; q++) r1[q] += (r2[q] + 1);
                    r1[19] += r2[19];
                    r1[18] += r2[18];
                    r1[18] += cn->score;
                    memmove(&g_subset.v[s2], &g_subset.v[s2 + 1], sizeof(okp_row) * (size_t) (g_subset.n - s2 - 1));
                    g_subset.n--;
                } else {
                    r1[part2] = (float) cn->cid2;
                    r1[19] += 1;
                    r1[18] += g_peaks_line.v[cn->cid2].score + cn->score;
                }
            } else if (found == 0 && pair_id < 18) {
                okp_row row;
                for (int q = 0; q < 20; q++) row.r[q] = -1;
                row.r[part1] = (float) cn->cid1;
                row.r[part2] = (float) cn->cid2;
                row.r[19] = 2;
                row.r[18] = g_peaks_line.v[cn->cid1].score + g_peaks_line.v[cn->cid2].score + cn->score;
                rowvec_push(&g_subset, row);
            }
        }
    }

    /* prune, pafprocess.cpp:187-191 */
    for (int i = g_subset.n - 1; i >= 0; i--) {
        const float *r = g_subset.v[i].r;
        if (r[19] < (float) OKP_THRESH_PART_CNT || r[18] / r[19] < OKP_THRESH_HUMAN_SCORE) {
            memmove(&g_subset.v[i], &g_subset.v[i + 1], sizeof(okp_row) * (size_t) (g_subset.n - i - 1));
            g_subset.n--;
        }
    }
    return 0;
}

/* getters, pafprocess.cpp:196-218 */
int okp_get_num_humans(void) { return g_subset.n; }
int okp_get_part_cid(int human_id, int part_id) { return (int) g_subset.v[human_id].r[part_id]; }
float okp_get_score(int human_id) { return g_subset.v[human_id].r[18] / g_subset.v[human_id].r[19]; }
int okp_get_part_x(int cid) { return g_peaks_line.v[cid].x; }
int okp_get_part_y(int cid) { return g_peaks_line.v[cid].y; }
float okp_get_part_score(int cid) { return g_peaks_line.v[cid].score; }

/* whole-state accessors for bit-exact diffs */
int okp_subset_rows(void) { return g_subset.n; }
void okp_subset_copy(float *dst) { memcpy(dst, g_subset.v, sizeof(okp_row) * (size_t) g_subset.n); }
int okp_num_peaks(void) { return g_peaks_line.n; }
void okp_peaks_copy(int *x, int *y, float *score, int *id) {
    for (int i = 0; i < g_peaks_line.n; i++) {
        x[i] = g_peaks_line.v[i].x; y[i] = g_peaks_line.v[i].y;
        score[i] = g_peaks_line.v[i].score; id[i] = g_peaks_line.v[i].id;
    }
}
int okp_num_connections(int limb) { return g_conns[limb].n; }
void okp_connections_copy(int limb, int *cid1, int *cid2, float *score, int *pid1, int *pid2) {
    for (int i = 0; i < g_conns[limb].n; i++) {
        const okp_conn *c = &g_conns[limb].v[i];
        cid1[i] = c->cid1; cid2[i] = c->cid2; score[i] = c->score; pid1[i] = c->peak_id1; pid2[i] = c->peak_id2;
    }
}
int okp_num_candidates(int limb) { return g_cands[limb].n; }
void okp_candidates_copy(int limb, int *idx1, int *idx2, float *score) {
    for (int i = 0; i < g_cands[limb].n; i++) {
        idx1[i] = g_cands[limb].v[i].idx1; idx2[i] = g_cands[limb].v[i].idx2; score[i] = g_cands[limb].v[i].score;
    }
}
