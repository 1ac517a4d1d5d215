// lantern_b200 -- batched multi-query HNSW search (the hot path).
//
// Replaces, for a whole batch of queries at once, the reference's per-query
//   index_gt::search                U/include/usearch/index.hpp:2680-2730
//   search_for_one_  (greedy)       :3277-3316
//   search_to_find_in_base_ (beam)  :3400-3485
//   sorted_buffer_gt / max_heap_gt  :529-780     (top-ef list + candidate queue)
//   growing_hash_set_gt (visited)   :922-1040
//   metric_punned_t                 index_plugins.hpp:1340-1342
// in "exact-order mode": the same decision sequence per query as the reference (see walk.cuh), so on
// tie-free data the returned ids are those of the reference on the same graph, at any ef.
//
// Mapping to the machine: CTAs are persistent and pull queries from an atomic counter; one CTA (4 warps)
// walks one query at a time.  Per expanded node, warp 0 reads the adjacency line (128 B at M=16), filters
// it against the CTA's visited bitmap in HBM and compacts the unseen ids; every warp then streams its share
// of the neighbour rows HBM -> shared memory with bulk async copies and reduces the distances; warp 0
// finally replays the reference's sequential insertions.  No row is read twice; besides rows the only HBM
// traffic is one adjacency line per expansion and the bitmap words.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <type_traits>

#include "walk.cuh"

namespace lb200 {

namespace {

template <class W>
__global__ void __launch_bounds__(kWalkThreads) hnsw_search_kernel(const __grid_constant__ SearchLaunch p, const uint32_t R) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const WalkLayout lay = W::layout(p.g, R, p.L, p.g.M0 * p.expand);
    W w(p.g);
    w.init(smem_raw, lay, R, p.s);
    WalkSmem& sm = w.sm;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            sm.ctrl->item = (uint32_t)atomicAdd(&p.s.counters[0], 1ull);
        __syncthreads();
        const uint32_t q = sm.ctrl->item;
        if (q >= p.nq)
            break;
        w.load_query(q, p.queries + (size_t)q * p.query_stride);

        uint32_t cur = p.g.entry;
        float cur_d = w.measure_one(cur);
        w.greedy(cur, cur_d, p.g.max_level, 0);
        const uint32_t size = w.beam(0, cur, cur_d, p.L, kNoNeighbor, p.expand);

        // results: top is ascending; shrink(k); keys (index.hpp:2722-2723, 2426-2433)
        if (w.warp == 0) {
            const uint32_t found = min(size, p.k);
            for (uint32_t i = w.lane; i < p.k; i += 32) {
                uint64_t key = ~0ull;
                float d = INFINITY;
                if (i < found) {
                    key = __ldg(p.g.keys + (sm.top_i[i] & kIdMask));
                    d = sm.top_d[i];
                }
                p.out_keys[(size_t)q * p.k + i] = key;
                p.out_dists[(size_t)q * p.k + i] = d;
            }
            if (w.lane == 0 && p.out_counts)
                p.out_counts[q] = found;
        }
    }
    if (threadIdx.x == 0) {
        atomicAdd(&p.s.counters[1], (unsigned long long)w.st_dist);
        atomicAdd(&p.s.counters[2], (unsigned long long)w.st_pops);
        atomicAdd(&p.s.counters[3], (unsigned long long)w.st_hops);
        if (w.st_limbo_drop)
            atomicAdd(&p.s.counters[4], (unsigned long long)w.st_limbo_drop);
    }
}

template <class W> void launch_one(const SearchLaunch& p, uint32_t R, size_t smem, uint32_t grid, int threads, cudaStream_t stream) {
    auto kern = hnsw_search_kernel<W>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, stream>>>(p, R);
    LB_CUDA(cudaGetLastError());
    count_launch();
}

template <class W> int occupancy_one(size_t smem, int threads) {
    auto kern = hnsw_search_kernel<W>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = 0;
    LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, threads, smem));
    return blocks;
}

} // namespace

// Warps per query: 4.  Two warps per query (twice the CTAs per SM, half the warps waiting at each barrier) was measured for
// 768-byte rows in round 2 and changed nothing (2.48 vs 2.50 M q/s on cfg5t, profiles/r02_bench_cfg5t_w*.json): the narrow-row
// bound is the per-expansion latency chain itself, not the idle warps.  LB200_SEARCH_WARPS = 2 keeps the experiment runnable.
static int search_threads(const GraphView& g, bool pq) {
    (void)g;
    int warps = 4;
    if (const char* e = getenv("LB200_SEARCH_WARPS"))
        if (!pq && (atoi(e) == 2 || atoi(e) == 4))
            warps = atoi(e);
    return warps * 32;
}

static size_t search_smem(const GraphView& g, bool pq, uint32_t R, uint32_t L, uint32_t expand) {
    return pq ? walk_layout_pq(g.num_subvectors, g.pq_lut_width, pq_value_floats(g), L, g.M0 * expand).total
              : walk_layout(R, g.row_bytes, L, g.M0 * expand).total;
}

uint32_t search_max_ctas(int dist_mode, int scalar_kind, const GraphView& g, uint32_t L, bool pq, uint32_t expand) {
    const uint32_t R = pick_ring_slots(g.row_bytes);
    const int nq = pq ? 1 : pick_nq(g.row_bytes);
    if (nq < 0)
        throw CudaError("search: vectors wider than 8192 bytes are not supported");
    const size_t smem = search_smem(g, pq, R, L, expand);
    int occ = 0;
    const int threads = search_threads(g, pq);
    dispatch_walker(pq, dist_mode, scalar_kind, nq, [&](auto tag) { occ = occupancy_one<typename decltype(tag)::type>(smem, threads); });
    if (occ < 1)
        throw CudaError("search: kernel does not fit on an SM (ef/k or the pq table too large for shared memory)");
    return (uint32_t)occ * (uint32_t)device_sm_count();
}

void launch_search(int dist_mode, int scalar_kind, bool pq, const SearchLaunch& p, cudaStream_t stream) {
    const uint32_t R = pick_ring_slots(p.g.row_bytes);
    const int nq = pq ? 1 : pick_nq(p.g.row_bytes);
    const size_t smem = search_smem(p.g, pq, R, p.L, p.expand);
    const uint32_t grid = p.s.ctas < p.nq ? p.s.ctas : p.nq;
    const int threads = search_threads(p.g, pq);
    dispatch_walker(pq, dist_mode, scalar_kind, nq,
                    [&](auto tag) { launch_one<typename decltype(tag)::type>(p, R, smem, grid, threads, stream); });
}

} // namespace lb200
