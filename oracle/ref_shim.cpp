// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin extension of the *unmodified* reference (usearch fork vendored in
// /root/reference/lantern_hnsw/third_party/usearch) that is compiled, from the
// sources where they lie, into oracle/_ref/liboracle_usearch.so next to the
// reference's own C shim (c/lib.cpp).  It adds no algorithm: it only exposes
// things the reference C API hides but the parity tests / CPU baseline need:
//   * per-query work counters (index.hpp:2370-2374 `computed_distances`, `visited_members`)
//   * a multi-threaded batch search / batch add driver (one usearch thread context per
//     host core, static chunking, as cpp/bench.cpp:312-335 does)
//   * add with an explicit level (lib.cpp:367-374 `usearch_add_external` path)
//
// The handle type is the one lib.cpp creates (lib.cpp:23 index_dense_t).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <usearch/index.hpp>
#include <usearch/index_dense.hpp>
#include <usearch/index_plugins.hpp>
#include <usearch/lantern_storage.hpp>

extern "C" {
#include "usearch.h"
}

using namespace unum::usearch;
using namespace unum;

using index_dense_t = index_dense_gt<default_key_t, lantern_slot_t, lantern_internal_storage_t, 'G'>;

namespace {

template <typename fn_t> void run_chunked(size_t n, size_t threads, fn_t&& fn) {
    if (threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; ++i)
            fn(0, i);
        return;
    }
    std::vector<std::thread> pool;
    size_t chunk = (n + threads - 1) / threads;
    for (size_t t = 0; t < threads; ++t) {
        size_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi)
            break;
        pool.emplace_back([=, &fn] {
            for (size_t i = lo; i < hi; ++i)
                fn(t, i);
        });
    }
    for (auto& th : pool)
        th.join();
}

} // namespace

extern "C" {

// One search, returning the reference's own work counters.
size_t refx_search_stats(usearch_index_t index, float const* query, size_t k, uint64_t* keys, float* dists,
                         uint64_t* computed_distances, uint64_t* visited_members) {
    auto* idx = reinterpret_cast<index_dense_t*>(index);
    auto r = idx->search(query, k, 0);
    if (computed_distances)
        *computed_distances = r.computed_distances;
    if (visited_members)
        *visited_members = r.visited_members;
    return r.dump_to(keys, dists);
}

// Batch search over `threads` usearch thread contexts. `queries` row-major, `stride` bytes.
// kind: 1 = f32, 5 = b1 (usearch_scalar_kind_t).  counts[q] = found; stats accumulate.
void refx_search_batch(usearch_index_t index, void const* queries, size_t nq, size_t stride, int kind, size_t k,
                       size_t threads, uint64_t* keys, float* dists, uint64_t* counts, uint64_t* sum_computed,
                       uint64_t* sum_visited) {
    auto* idx = reinterpret_cast<index_dense_t*>(index);
    std::atomic<uint64_t> comp{0}, vis{0};
    run_chunked(nq, threads, [&](size_t t, size_t q) {
        char const* qp = (char const*)queries + q * stride;
        auto r = kind == usearch_scalar_b1_k ? idx->search((b1x8_t const*)qp, k, t)
                                             : idx->search((f32_t const*)qp, k, t);
        comp += r.computed_distances;
        vis += r.visited_members;
        size_t found = r.dump_to(keys + q * k, dists + q * k);
        if (counts)
            counts[q] = found;
    });
    if (sum_computed)
        *sum_computed = comp.load();
    if (sum_visited)
        *sum_visited = vis.load();
}

// Batch add over `threads` contexts (what lantern_cli's indexer threads do, server.rs:328-359).
// Returns total computed distances.
uint64_t refx_add_batch(usearch_index_t index, uint64_t const* keys, void const* vectors, size_t n, size_t stride,
                        int kind, size_t threads) {
    auto* idx = reinterpret_cast<index_dense_t*>(index);
    std::atomic<uint64_t> comp{0};
    // first element alone, so that the entry point exists before the threads start
    size_t start = 0;
    if (idx->size() == 0 && n) {
        char const* vp = (char const*)vectors;
        auto r = kind == usearch_scalar_b1_k ? idx->add(keys[0], (b1x8_t const*)vp, 0)
                                             : idx->add(keys[0], (f32_t const*)vp, 0);
        comp += r.computed_distances;
        start = 1;
    }
    run_chunked(n - start, threads, [&](size_t t, size_t i) {
        i += start;
        char const* vp = (char const*)vectors + i * stride;
        auto r = kind == usearch_scalar_b1_k ? idx->add(keys[i], (b1x8_t const*)vp, t)
                                             : idx->add(keys[i], (f32_t const*)vp, t);
        comp += r.computed_distances;
    });
    return comp.load();
}

// Add with an explicit level on thread context 0 (deterministic, single-threaded).
uint64_t refx_add_level(usearch_index_t index, uint64_t key, void const* vector, int kind, int level) {
    auto* idx = reinterpret_cast<index_dense_t*>(index);
    auto r = kind == usearch_scalar_b1_k ? idx->add(key, (b1x8_t const*)vector, 0, true, (level_t)level)
                                         : idx->add(key, (f32_t const*)vector, 0, true, (level_t)level);
    return r.computed_distances;
}

size_t refx_hardware_threads() { return std::thread::hardware_concurrency(); }
}
