/* Plain-C caller of the libcrs C ABI (include/crs.h): index a few rows, run one exact search.
 *
 *   gcc -std=c99 -Iinclude examples/crs_example.c -o crs_example \
 *       -Lcompressed_rag_suite_b200 -lcrs -Wl,-rpath,$PWD/compressed_rag_suite_b200 -lm
 *
 * This is what a non-Python host of the reference's retrieval path would write in place of
 * collection.add / collection.query (reference rag/indexing.py:114-119,171-176).  Without an
 * sm_100 GPU crs_index_create fails with CRS_ECUDA and the program says so: the library has
 * no CPU implementation. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "crs.h"

#define N 4096
#define DIM 384
#define K 5

static float frand(unsigned* s) {
    *s = *s * 1664525u + 1013904223u;
    return (float)((*s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}

int main(void) {
    crs_index* ix = NULL;
    int rc = crs_index_create(&ix, DIM, CRS_F16, CRS_COSINE, /*device*/ 0, /*row_base*/ 0, /*reserve*/ N);
    if (rc != CRS_OK) {
        fprintf(stderr, "crs_index_create failed (status %d): %s\n", rc, crs_last_error());
        return 2;
    }
    float* rows = (float*)malloc(sizeof(float) * N * DIM);
    unsigned seed = 7;
    for (long i = 0; i < (long)N * DIM; ++i) rows[i] = frand(&seed);
    rc = crs_index_add(ix, rows, N, CRS_F32);                 /* normalised + stored as fp16 on the GPU */
    if (rc != CRS_OK) { fprintf(stderr, "crs_index_add: %s\n", crs_last_error()); return 3; }

    uint32_t ids[K];
    float sims[K];
    int32_t count = 0;
    rc = crs_index_search(ix, rows + 123 * DIM, 1, K, -INFINITY, ids, sims, &count);   /* query = row 123 */
    if (rc != CRS_OK) { fprintf(stderr, "crs_index_search: %s\n", crs_last_error()); return 4; }
    for (int i = 0; i < count; ++i)
        printf("rank %d: row %u  cosine %.6f  chroma distance %.6f\n", i, ids[i], sims[i], 1.0 - (double)sims[i]);
    int ok = count == K && ids[0] == 123u && sims[0] > 0.999f;
    crs_index_destroy(ix);
    free(rows);
    printf(ok ? "ok\n" : "UNEXPECTED RESULT\n");
    return ok ? 0 : 1;
}
