/* Source-level drop-in check: this file is compiled against the REFERENCE's own header (usearch.h, found under
 * /root/reference at build time, never copied) and linked against liblantern_b200.so.  It plays the call sequence of
 * lantern_hnsw/src/hnsw/build.c (init, reserve, add, save_buffer) and scan.c (search_ef) on the small_world cube.
 * Exit codes: 0 = ran on a GPU and the nearest neighbour is right; 3 = no CUDA device (the library said so through the
 * usearch error convention); anything else = failure. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "usearch.h"

/* build.c:512-517 fills the retriever fields before usearch_init although they are never called on the build path */
static void* fake_retriever(void* ctx, unsigned long long id) {
    (void)ctx, (void)id;
    abort(); /* the engine must never call back into page storage */
}

int main(void) {
    usearch_init_options_t opts;
    memset(&opts, 0, sizeof(opts));
    opts.retriever_ctx = &opts;
    opts.retriever = (usearch_node_retriever_t)fake_retriever;
    opts.retriever_mut = (usearch_node_retriever_t)fake_retriever;
    opts.metric_kind = usearch_metric_l2sq_k;
    opts.quantization = usearch_scalar_f32_k;
    opts.dimensions = 3;
    opts.connectivity = 2;
    opts.expansion_add = 128;
    opts.expansion_search = 4;
    opts.num_threads = 1;
    usearch_error_t error = NULL;
    usearch_index_t idx = usearch_init(&opts, NULL, &error);
    if (error) {
        printf("usearch_init: %s\n", error);
        return strstr(error, "CUDA device unavailable") ? 3 : 1;
    }
    usearch_reserve(idx, 8, &error);
    if (error)
        return 1;
    for (int i = 0; i < 8; ++i) {
        float v[3] = {(float)((i >> 2) & 1), (float)((i >> 1) & 1), (float)(i & 1)};
        usearch_add(idx, (usearch_key_t)(100 + i), v, usearch_scalar_f32_k, &error);
        if (error) {
            printf("usearch_add: %s\n", error);
            return 1;
        }
    }
    if (usearch_size(idx, &error) != 8 || usearch_dimensions(idx, &error) != 3 || usearch_connectivity(idx, &error) != 2)
        return 1;
    float q[3] = {0.f, 1.f, 0.f};
    usearch_key_t keys[8];
    usearch_distance_t dists[8];
    size_t found = usearch_search_ef(idx, q, usearch_scalar_f32_k, 8, 0, false, keys, dists, &error);
    if (error || found != 8 || keys[0] != 102 || dists[0] != 0.f || dists[7] != 3.f) {
        printf("search: %s found=%zu key0=%llu\n", error ? error : "", found, (unsigned long long)keys[0]);
        return 1;
    }
    size_t len = usearch_serialized_length(idx, &error);
    char* buf = (char*)malloc(len);
    usearch_save_buffer(idx, buf, len, &error);
    if (error || memcmp(buf, "usearch", 7) != 0 || usearch_header_get_entry_slot(buf) >= 8)
        return 1;
    char header[136]; /* insert.c persists this into the index header page after every insert */
    usearch_update_header(idx, header, &error);
    if (error || memcmp(header, buf, sizeof(header)) != 0)
        return 1;
    if (usearch_count(idx, 102, &error) != 1 || usearch_contains(idx, 5, &error) || error)
        return 1;
    usearch_view_mem_lazy(idx, header, &error); /* page storage: must refuse, not crash */
    if (!error)
        return 1;
    error = NULL;
    usearch_index_metadata_t meta = usearch_index_metadata(idx, &error);
    if (meta.neighbors_bytes != 16 || meta.neighbors_base_bytes != 28) /* 4+6M / 4+12M for M=2 */
        return 1;
    free(buf);
    usearch_free(idx, &error);
    printf("drop-in ok: nearest key %llu\n", (unsigned long long)keys[0]);
    return 0;
}
