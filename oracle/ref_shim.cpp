/* oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, never part of the product path.
 *
 * Gives the UNMODIFIED reference implementation C linkage so ctypes can call
 * it (the reference binds it with SWIG, which is absent here and adds no
 * arithmetic: lib/pafprocess/pafprocess.i:1-15).  The reference source is
 * #included from where it lies under /root/reference (never copied into this
 * repository); the Makefile passes -I$(REF)/lib/pafprocess.  Output goes to
 * oracle/_ref/libpaf_ref.so only.
 *
 * Reference interface wrapped: lib/pafprocess/pafprocess.h:53-59.
 * The two extra accessors read the reference's own globals `subset` and
 * `peak_infos_line` (lib/pafprocess/pafprocess.cpp:12-13) so tests can diff
 * whole rows bit-for-bit instead of going through the int-truncating getters.
 */
#include "pafprocess.cpp"

extern "C" {

int ref_process_paf(int p1, int p2, int p3, float *peaks, int h1, int h2, int h3, float *heat,
                    int f1, int f2, int f3, float *paf) {
    return process_paf(p1, p2, p3, peaks, h1, h2, h3, heat, f1, f2, f3, paf);
}
int ref_get_num_humans() { return get_num_humans(); }
int ref_get_part_cid(int human_id, int part_id) { return get_part_cid(human_id, part_id); }
float ref_get_score(int human_id) { return get_score(human_id); }
int ref_get_part_x(int cid) { return get_part_x(cid); }
int ref_get_part_y(int cid) { return get_part_y(cid); }
float ref_get_part_score(int cid) { return get_part_score(cid); }

/* whole-row accessors (bit-exact diffs) */
int ref_subset_rows() { return (int) subset.size(); }
void ref_subset_copy(float *dst /* [rows][20] */) {
    for (size_t i = 0; i < subset.size(); i++)
        for (int k = 0; k < 20; k++) dst[i * 20 + k] = subset[i][k];
}
int ref_num_peaks() { return (int) peak_infos_line.size(); }
void ref_peaks_copy(int *x, int *y, float *score, int *id) {
    for (size_t i = 0; i < peak_infos_line.size(); i++) {
        x[i] = peak_infos_line[i].x;
        y[i] = peak_infos_line[i].y;
        score[i] = peak_infos_line[i].score;
        id[i] = peak_infos_line[i].id;
    }
}

/* std::sort exactly as the reference calls it (pafprocess.cpp:97: same element type, same
 * comparator), on caller-supplied scores; `tag` rides in idx1 so the tie permutation is visible. */
void ref_std_sort(int n, float *score, int *tag) {
    vector<ConnectionCandidate> v(n);
    for (int i = 0; i < n; i++) { v[i].idx1 = tag[i]; v[i].idx2 = 0; v[i].score = score[i]; v[i].etc = 0; }
    sort(v.begin(), v.end(), comp_candidate);
    for (int i = 0; i < n; i++) { score[i] = v[i].score; tag[i] = v[i].idx1; }
}

}  // extern "C"
