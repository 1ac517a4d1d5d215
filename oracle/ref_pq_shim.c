/* TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference k-means (product_quantization.c, compiled from
 * /root/reference by `make refpq`) from plain arrays: builds the float** dataset it expects, scripts the PRNG so that
 * the initial centres of subvector s are rows init_rows[s][0..k), and lays the result out as the codebook tape
 * float[num_centroids][dim] that Lantern's load_pq_codebook produces (pqtable.c:194-333) and usearch_init consumes. */
#include <postgres.h>

#include <string.h>

#include "product_quantization.h"

static const uint32_t* g_script;
static size_t g_script_len, g_script_pos;

long oracle_scripted_random(void) {
    if (g_script_pos >= g_script_len)
        abort(); /* a duplicate row in the script made get_random_tid draw again: the test must pass distinct rows */
    return (long)g_script[g_script_pos++];
}

/* returns 0; codebook[c * dim + s * subdim + j] = centroid c of subvector s.  dim must be divisible by nsub. */
int refpq_train(const float* data, uint32_t n, uint32_t dim, uint32_t nsub, uint32_t ncent, int metric, uint32_t iter,
                const uint32_t* init_rows, float* codebook) {
    float** rows = (float**)malloc(sizeof(float*) * n);
    uint32_t i, s, c;
    for (i = 0; i < n; ++i)
        rows[i] = (float*)(data + (size_t)i * dim);
    g_script = init_rows, g_script_len = (size_t)nsub * ncent, g_script_pos = 0;
    PQCodebook** books = product_quantization(ncent, nsub, rows, n, dim, (usearch_metric_kind_t)metric, iter);
    const uint32_t sd = dim / nsub;
    for (s = 0; s < nsub; ++s)
        for (c = 0; c < ncent; ++c)
            memcpy(codebook + (size_t)c * dim + (size_t)s * sd, books[s]->centroids[c], sizeof(float) * sd);
    free(rows);
    return 0;
}
