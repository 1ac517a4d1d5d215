// lantern_b200 -- HNSW construction on the GPU.
//
// Replaces the reference's insertion path
//   index_gt::add                       U/include/usearch/index.hpp:2479-2564
//   connect_node_across_levels_         :3119-3136
//   search_to_insert_                   :3324-3392   (Walker::beam on level l, width = expansion_add)
//   connect_new_node_ / refine_         :3139-3160, :3515-3561   (Walker::refine to `connectivity` on EVERY level)
//   reconnect_neighbor_nodes_           :3163-3206   (reverse links; re-prune to M / M0 when a list is full)
//   choose_random_level_                :3208-3212   (same generator: minstd_rand0 + generate_canonical<double,53>)
// for BATCHES of vectors.  A batch is inserted in two kernels:
//   1. build_insert_kernel : one CTA per new node: greedy descent, efc-wide beam per level, heuristic neighbour
//      selection, writes the node's own lists and emits one reverse-link request per selected neighbour;
//   2. (radix sort of the requests by (level, target), segment heads)
//   3. build_reverse_kernel: one CTA per (level, target) segment applies the requests in insertion order: append
//      while the list has room, otherwise re-prune {new} + list with the same heuristic, distances measured from
//      the target (index.hpp:3194-3198).
// New nodes of one batch do not see each other during step 1 (they are linked to older nodes only).  With batch
// size 1 ("exact-order build") the procedure is the reference's sequential insertion: same levels, same lists,
// byte-identical file on data whose fp32 sums do not depend on summation order.  A node whose level exceeds the
// current top level always forms its own batch so that entry-point updates stay sequential.
#include <cuda_runtime.h>

#include <cub/cub.cuh>
#include <type_traits>
#include <vector>

#include "walk.cuh"

namespace lb200 {

namespace {

struct BuildLaunch {
    GraphView g;
    SearchScratch s;
    uint32_t* adj0;      // mutable views of g.adj0 / g.upper_adj
    uint32_t* upper_adj;
    uint32_t u0, count;        // batch = nodes [u0, u0 + count)
    const int32_t* new_levels; // [count]
    const uint8_t* values;     // value (query side) of batch item i at values + i * value_stride
    size_t value_stride;
    uint32_t efc;
    uint32_t top_cap;
    uint32_t req_stride; // requests per batch item = M * (max_level + 1)
    unsigned long long* req_keys; // (level << 32) | target, ~0 = empty
    unsigned long long* req_vals; // (float bits of d(new, target) << 32) | new
    // reverse phase
    const unsigned long long* sorted_keys;
    const unsigned long long* sorted_vals;
    const uint32_t* seg_start;
    const uint32_t* nseg;
    uint32_t total_reqs;
};

template <class W>
__global__ void __launch_bounds__(kWalkThreads) build_insert_kernel(const __grid_constant__ BuildLaunch p, const uint32_t R) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const WalkLayout lay = W::layout(p.g, R, p.top_cap, p.g.M0 + 1);
    W w(p.g);
    w.init(smem_raw, lay, R, p.s);
    WalkSmem& sm = w.sm;
    const uint32_t M = p.g.M, M0 = p.g.M0;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            sm.ctrl->item = (uint32_t)atomicAdd(&p.s.counters[0], 1ull);
        __syncthreads();
        const uint32_t item = sm.ctrl->item;
        if (item >= p.count)
            break;
        const uint32_t u = p.u0 + item;
        const int lu = p.new_levels[item];
        const uint8_t* urow = p.values + (size_t)item * p.value_stride;
        w.load_value(urow);

        uint32_t cur = p.g.entry;
        float cur_d = w.measure_one(cur);
        w.greedy(cur, cur_d, p.g.max_level, lu); // levels max_level .. lu+1
        for (int level = min(lu, p.g.max_level); level >= 0; --level) {
            const uint32_t size = w.beam(level, cur, cur_d, p.efc, u);
            const uint32_t view = w.refine(size, M); // connect_new_node_: `connectivity` on every level (:3149)
            const uint32_t width = level ? M : M0;
            uint32_t* list = level ? p.upper_adj + ((size_t)p.g.upper_ref[u] + (level - 1)) * M : p.adj0 + (size_t)u * M0;
            for (uint32_t j = threadIdx.x; j < width; j += kWalkThreads)
                list[j] = j < view ? (sm.top_i[j] & kIdMask) : kNoNeighbor;
            for (uint32_t j = threadIdx.x; j < view; j += kWalkThreads) {
                const size_t slot = (size_t)item * p.req_stride + (size_t)level * M + j;
                const uint32_t v = sm.top_i[j] & kIdMask;
                p.req_keys[slot] = ((unsigned long long)(uint32_t)level << 32) | v;
                p.req_vals[slot] = ((unsigned long long)__float_as_uint(sm.top_d[j]) << 32) | u;
            }
            cur = sm.top_i[0] & kIdMask, cur_d = sm.top_d[0]; // closest_slot = new_neighbors[0] (:3159)
            __syncthreads();
            if (level > 0)
                w.load_value(urow); // refine() used the value registers for candidate rows
        }
    }
    if (threadIdx.x == 0) {
        atomicAdd(&p.s.counters[1], (unsigned long long)w.st_dist);
        atomicAdd(&p.s.counters[2], (unsigned long long)w.st_pops);
        atomicAdd(&p.s.counters[3], (unsigned long long)w.st_hops);
    }
}

__global__ void segment_heads_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ seg_start,
                                     uint32_t* __restrict__ nseg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const unsigned long long k = keys[i];
    if (k == ~0ull)
        return;
    if (i == 0 || keys[i - 1] != k)
        seg_start[atomicAdd(nseg, 1u)] = i;
}

template <class W>
__global__ void __launch_bounds__(kWalkThreads) build_reverse_kernel(const __grid_constant__ BuildLaunch p, const uint32_t R) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t M = p.g.M, M0 = p.g.M0;
    const WalkLayout lay = W::layout(p.g, R, M0 + 2, M0 + 1);
    W w(p.g);
    w.init(smem_raw, lay, R, p.s);
    WalkSmem& sm = w.sm;
    uint32_t* lst_i = reinterpret_cast<uint32_t*>(smem_raw + lay.total);
    float* lst_d = reinterpret_cast<float*>(lst_i + M0);
    const uint32_t nseg = *p.nseg;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            sm.ctrl->item = (uint32_t)atomicAdd(&p.s.counters[0], 1ull);
        __syncthreads();
        const uint32_t seg = sm.ctrl->item;
        if (seg >= nseg)
            break;
        const uint32_t s0 = p.seg_start[seg];
        const unsigned long long key = p.sorted_keys[s0];
        const int level = (int)(key >> 32);
        const uint32_t v = (uint32_t)key;
        const uint32_t cmax = level ? M : M0;
        uint32_t* list = level ? p.upper_adj + ((size_t)p.g.upper_ref[v] + (level - 1)) * M : p.adj0 + (size_t)v * M0;

        for (uint32_t j = threadIdx.x; j < cmax; j += kWalkThreads)
            lst_i[j] = list[j];
        __syncthreads();
        uint32_t cnt = 0; // leading valid entries (lists are compact)
        while (cnt < cmax && lst_i[cnt] != kNoNeighbor)
            ++cnt;
        bool have_d = false;

        for (uint32_t r = s0; r < p.total_reqs && p.sorted_keys[r] == key; ++r) {
            const unsigned long long val = p.sorted_vals[r];
            const uint32_t u = (uint32_t)val;
            const float d_uv = __uint_as_float((uint32_t)(val >> 32));
            __syncthreads();
            if (cnt < cmax) { // index.hpp:3186-3189
                if (threadIdx.x == 0)
                    lst_i[cnt] = u, lst_d[cnt] = d_uv;
                cnt++;
                continue;
            }
            if (!have_d) { // distances target -> each current neighbour (index.hpp:3196-3198)
                w.load_node(v);
                for (uint32_t j = threadIdx.x; j < cnt; j += kWalkThreads)
                    sm.cand_id[j] = lst_i[j];
                __syncthreads();
                w.eval(cnt);
                __syncthreads();
                for (uint32_t j = threadIdx.x; j < cnt; j += kWalkThreads)
                    lst_d[j] = sm.cand_d[j];
                w.st_dist += cnt;
                have_d = true;
                __syncthreads();
            }
            if (w.warp == 0) { // top = sorted {new} + successors, inserted in this order (insert_reserved, :3192-3198)
                uint32_t size = 0, cursor = 0;
                top_insert(sm.top_d, sm.top_i, size, cursor, cnt + 2, d_uv, u, w.lane);
                for (uint32_t j = 0; j < cnt; ++j)
                    top_insert(sm.top_d, sm.top_i, size, cursor, cnt + 2, lst_d[j], lst_i[j], w.lane);
            }
            __syncthreads();
            const uint32_t view = w.refine(cnt + 1, cmax);
            for (uint32_t j = threadIdx.x; j < view; j += kWalkThreads)
                lst_i[j] = sm.top_i[j] & kIdMask, lst_d[j] = sm.top_d[j];
            cnt = view;
        }
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < cmax; j += kWalkThreads)
            list[j] = j < cnt ? lst_i[j] : kNoNeighbor;
    }
    if (threadIdx.x == 0)
        atomicAdd(&p.s.counters[1], (unsigned long long)w.st_dist);
}

// reference level generator (index.hpp:3208-3212 with libstdc++'s minstd_rand0 / generate_canonical<double,53>)
int draw_level(uint32_t& state, size_t M) {
    auto next = [&]() {
        state = (uint32_t)(((uint64_t)state * 16807ull) % 2147483647ull);
        return state;
    };
    const double Rr = 2147483646.0;
    double sum = (double)(next() - 1u);
    sum += (double)(next() - 1u) * Rr;
    double u = sum / (Rr * Rr);
    if (u >= 1.0)
        u = nextafter(1.0, 0.0);
    return (int)(int16_t)(-log(u) * (1.0 / log((double)M)));
}

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void ensure(size_t want) {
        if (want <= n)
            return;
        if (p)
            LB_CUDA(cudaFree(p));
        p = nullptr;
        n = want + want / 2;
        LB_CUDA(cudaMalloc(&p, n * sizeof(T)));
    }
    ~DevBuf() { cudaFree(p); }
};

} // namespace

static void build_pending_impl(Index& idx);

// Bookkeeping is transactional: if anything below throws (out of memory, launch failure) the level generator, the level /
// list counters and the node count are put back, so that a retry draws the same levels and inserts the same nodes again
// instead of double-inserting (pending_n_ is only cleared on success).
void build_pending(Index& idx) {
    if (!idx.pending_n_)
        return;
    const uint32_t rng0 = idx.level_rng_;
    const size_t n0 = idx.n_, lists0 = idx.upper_lists_;
    const int32_t max_level0 = idx.max_level_;
    const uint32_t entry0 = idx.entry_;
    try {
        build_pending_impl(idx);
    } catch (...) {
        idx.level_rng_ = rng0, idx.n_ = n0, idx.upper_lists_ = lists0, idx.max_level_ = max_level0, idx.entry_ = entry0;
        idx.h_levels_.resize(n0);
        (void)cudaGetLastError();
        throw;
    }
}

static void build_pending_impl(Index& idx) {
    const size_t P = idx.pending_n_;
    if (!P)
        return;
    const IndexConfig& cfg = idx.cfg_;
    const size_t n0 = idx.n_;
    const uint32_t M = (uint32_t)cfg.M, M0 = (uint32_t)cfg.M0;
    cudaStream_t stream = 0;

    // ---- levels + upper-list allocation for the new nodes ----
    idx.h_levels_.resize(n0 + P);
    std::vector<int32_t> new_levels(P);
    std::vector<uint32_t> new_upper_ref(P, kNoNeighbor);
    size_t lists = idx.upper_lists_;
    for (size_t i = 0; i < P; ++i) {
        int l = draw_level(idx.level_rng_, cfg.M);
        new_levels[i] = l;
        idx.h_levels_[n0 + i] = (int16_t)l;
        if (l > 0) {
            new_upper_ref[i] = (uint32_t)lists;
            lists += (size_t)l;
        }
    }
    idx.alloc_upper(lists + 1);
    idx.upper_lists_ = lists;
    DevBuf<int32_t> d_levels;
    d_levels.ensure(P);
    LB_CUDA(cudaMemcpyAsync(d_levels.p, new_levels.data(), P * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    LB_CUDA(cudaMemcpyAsync(idx.d_upper_ref_ + n0, new_upper_ref.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    LB_CUDA(cudaMemcpyAsync(idx.d_keys_ + n0, idx.h_keys_.data() + n0, P * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));

    // ---- kernel geometry ----
    const uint32_t row_bytes = (uint32_t)idx.row_bytes_;
    const uint32_t R = pick_ring_slots(row_bytes);
    const int nq = cfg.pq ? 1 : pick_nq(row_bytes);
    if (nq < 0)
        throw CudaError("build: vectors wider than 8192 bytes are not supported");
    const uint32_t top_cap = (uint32_t)std::max<size_t>(cfg.efc, M0 + 2);
    const GraphView gv0 = idx.view();
    const size_t smem_ins = cfg.pq ? walk_layout_pq(gv0.num_subvectors, gv0.pq_lut_width, pq_value_floats(gv0), top_cap, M0 + 1).total
                                   : walk_layout(R, row_bytes, top_cap, M0 + 1).total;
    const size_t smem_rev = (cfg.pq ? walk_layout_pq(gv0.num_subvectors, gv0.pq_lut_width, pq_value_floats(gv0), M0 + 2, M0 + 1).total
                                    : walk_layout(R, row_bytes, M0 + 2, M0 + 1).total) + (size_t)M0 * 8;
    int occ_ins = 0, occ_rev = 0;
    dispatch_walker(cfg.pq, idx.dist_mode_, cfg.scalar_kind, nq, [&](auto tag) {
        using W = typename decltype(tag)::type;
        auto k1 = build_insert_kernel<W>;
        auto k3 = build_reverse_kernel<W>;
        LB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ins));
        LB_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rev));
        LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ins, k1, kWalkThreads, smem_ins));
        LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_rev, k3, kWalkThreads, smem_rev));
    });
    if (occ_ins < 1 || occ_rev < 1)
        throw CudaError("build: kernel does not fit on an SM (expansion_add too large for shared memory?)");
    const uint32_t sms = (uint32_t)device_sm_count();
    const uint32_t max_ctas = std::max(occ_ins, occ_rev) * sms;
    idx.ensure_scratch(max_ctas);

    const size_t batch_cap = idx.build_batch_ ? idx.build_batch_ : (size_t)occ_ins * sms;
    DevBuf<unsigned long long> req_keys, req_vals, srt_keys, srt_vals;
    DevBuf<uint32_t> seg_start, nseg;
    DevBuf<uint8_t> cub_tmp;
    nseg.ensure(1);

    // work counters of this build (SURVEY 8d build metric: sum of computed_distances(add) x row bytes / time)
    LB_CUDA(cudaMemsetAsync(idx.scratch_.counters, 0, 8 * sizeof(unsigned long long), stream));
    LB_CUDA(cudaEventRecord(idx.ev0_, stream));

    size_t pos = 0;
    if (n0 == 0) { // first node: entry point, no links (index.hpp:2538-2543)
        idx.entry_ = 0;
        idx.max_level_ = new_levels[0];
        idx.n_ = 1;
        pos = 1;
    }
    while (pos < P) {
        const size_t visible = idx.n_;
        size_t bsz = 1;
        if (new_levels[pos] <= idx.max_level_) {
            const size_t ramp = std::max<size_t>(1, visible / idx.build_ratio_);
            bsz = std::min(std::min(P - pos, batch_cap), ramp);
            for (size_t i = 1; i < bsz; ++i)
                if (new_levels[pos + i] > idx.max_level_) {
                    bsz = i;
                    break;
                }
        }
        const uint32_t u0 = (uint32_t)(n0 + pos);
        const uint32_t levels_here = (uint32_t)idx.max_level_ + 1;
        const uint32_t req_stride = M * levels_here;
        const size_t nreq = bsz * req_stride;
        req_keys.ensure(nreq), req_vals.ensure(nreq), srt_keys.ensure(nreq), srt_vals.ensure(nreq), seg_start.ensure(nreq);
        LB_CUDA(cudaMemsetAsync(req_keys.p, 0xFF, nreq * sizeof(unsigned long long), stream));
        LB_CUDA(cudaMemsetAsync(idx.scratch_.counters, 0, sizeof(unsigned long long), stream));
        LB_CUDA(cudaMemsetAsync(nseg.p, 0, sizeof(uint32_t), stream));

        BuildLaunch p{};
        p.g = idx.view();
        p.s = idx.scratch_;
        p.adj0 = idx.d_adj0_, p.upper_adj = idx.d_upper_adj_;
        p.u0 = u0, p.count = (uint32_t)bsz;
        p.new_levels = d_levels.p + pos;
        if (cfg.pq) { // value = the raw f32 vector; stored side = codes
            p.value_stride = round_up(cfg.dims * 4, 16);
            p.values = (const uint8_t*)idx.d_pending_raw_ + pos * p.value_stride;
        } else {
            p.value_stride = row_bytes;
            p.values = idx.d_vectors_ + (size_t)u0 * row_bytes;
        }
        p.efc = (uint32_t)cfg.efc, p.top_cap = top_cap;
        p.req_stride = req_stride;
        p.req_keys = req_keys.p, p.req_vals = req_vals.p;
        p.sorted_keys = srt_keys.p, p.sorted_vals = srt_vals.p;
        p.seg_start = seg_start.p, p.nseg = nseg.p;
        p.total_reqs = (uint32_t)nreq;

        const uint32_t grid_ins = (uint32_t)std::min<size_t>(bsz, (size_t)occ_ins * sms);
        dispatch_walker(cfg.pq, idx.dist_mode_, cfg.scalar_kind, nq, [&](auto tag) {
            build_insert_kernel<typename decltype(tag)::type><<<grid_ins, kWalkThreads, smem_ins, stream>>>(p, R);
        });
        LB_CUDA(cudaGetLastError());
        count_launch();

        size_t tmp_bytes = 0;
        const int end_bit = 32 + 8; // levels < 256
        LB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, req_keys.p, srt_keys.p, req_vals.p, srt_vals.p, (int)nreq, 0,
                                                end_bit, stream));
        cub_tmp.ensure(tmp_bytes);
        // keys ~0 (empty) have all bits set and therefore sort last within the examined bits as well
        LB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, req_keys.p, srt_keys.p, req_vals.p, srt_vals.p, (int)nreq, 0,
                                                end_bit, stream));
        count_launch(2);
        segment_heads_kernel<<<(unsigned)((nreq + 255) / 256), 256, 0, stream>>>(srt_keys.p, (uint32_t)nreq, seg_start.p, nseg.p);
        LB_CUDA(cudaGetLastError());
        count_launch();
        LB_CUDA(cudaMemsetAsync(idx.scratch_.counters, 0, sizeof(unsigned long long), stream));
        const uint32_t grid_rev = (uint32_t)std::min<size_t>(nreq, (size_t)occ_rev * sms);
        dispatch_walker(cfg.pq, idx.dist_mode_, cfg.scalar_kind, nq, [&](auto tag) {
            build_reverse_kernel<typename decltype(tag)::type><<<grid_rev, kWalkThreads, smem_rev, stream>>>(p, R);
        });
        LB_CUDA(cudaGetLastError());
        count_launch();

        idx.n_ += bsz;
        if (new_levels[pos] > idx.max_level_) { // index.hpp:2558-2562
            idx.entry_ = u0;
            idx.max_level_ = new_levels[pos];
        }
        pos += bsz;
    }
    LB_CUDA(cudaEventRecord(idx.ev1_, stream));
    LB_CUDA(cudaStreamSynchronize(stream));
    idx.pending_n_ = 0;
    {
        unsigned long long c[4] = {0, 0, 0, 0};
        LB_CUDA(cudaMemcpy(c, idx.scratch_.counters, sizeof(c), cudaMemcpyDeviceToHost));
        float ms = 0.f;
        LB_CUDA(cudaEventElapsedTime(&ms, idx.ev0_, idx.ev1_));
        idx.last_build_ms_ = ms, idx.last_build_dist_ = c[1], idx.last_build_n_ = P;
        idx.last_nq_ = 0; // the events now bracket a build, not a search
    }
    if (idx.d_pending_raw_) { // raw rows are only needed while inserting
        LB_CUDA(cudaFree(idx.d_pending_raw_));
        idx.d_pending_raw_ = nullptr, idx.pending_raw_cap_ = 0;
    }
}

} // namespace lb200
