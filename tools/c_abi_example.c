/* Plain C99 caller of libekpose_b200.so: the reference operator surface on host pointers (the seven symbols of
 * lib/pafprocess/pafprocess.h:53-59) and a batched run through the handle API.  Built by tests/test_abi.py as a
 * header / link check (no GPU needed to build):
 *   gcc -std=c99 -Wall -Iinclude tools/c_abi_example.c -Ltorch_ekpose_b200 -lekpose_b200 -Wl,-rpath,$PWD/torch_ekpose_b200 -o c_abi_example
 * Run on a GPU box it prints the known answer of SURVEY.md Appendix A.6: 1 human, score 1.5. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ekpose_b200.h"

int main(void) {
    enum { H = 64, W = 64, C = 38 };
    float *paf = (float *) calloc((size_t) H * W * C, sizeof(float));
    float peaks[4][5] = {{10, 10, .9f, 0, 1}, {20, 10, .8f, 0, 2}, {30, 10, .7f, 0, 3}, {40, 10, .6f, 0, 4}};
    int y, x, rc;
    if (!paf) return 2;
    for (y = 0; y < H; y++)
        for (x = 0; x < W; x++) {
            paf[(y * W + x) * C + 12] = 1.f;  /* limbs 0, 2, 3: x channels 12, 14, 16 (pafprocess.h:16-19) */
            paf[(y * W + x) * C + 14] = 1.f;
            paf[(y * W + x) * C + 16] = 1.f;
        }
    rc = process_paf(1, 4, 5, &peaks[0][0], H, W, 19, NULL, H, W, C, paf);
    if (rc != EKP_OK) {
        fprintf(stderr, "process_paf: %d (%s)\n", rc, ekp_last_error());
        free(paf);
        return rc == EKP_ERR_CUDA ? 3 : 1;  /* 3: no GPU here -- there is no CPU fallback */
    }
    printf("%s: humans %d, score %.3f, parts 1..4 -> cids %d %d %d %d, x of cid 3 = %d\n", ekp_version(), get_num_humans(),
           get_score(0), get_part_cid(0, 1), get_part_cid(0, 2), get_part_cid(0, 3), get_part_cid(0, 4), get_part_x(3));
    free(paf);
    return get_num_humans() == 1 ? 0 : 1;
}
