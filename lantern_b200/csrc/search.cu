// lantern_b200 -- batched multi-query HNSW search (the hot path).
//
// Replaces, for a whole batch of queries at once, the reference's per-query
//   index_gt::search                U/include/usearch/index.hpp:2680-2730
//   search_for_one_  (greedy)       :3277-3316
//   search_to_find_in_base_ (beam)  :3400-3485
//   sorted_buffer_gt / max_heap_gt  :529-780     (top-ef list + candidate queue)
//   growing_hash_set_gt (visited)   :922-1040
//   metric_punned_t                 index_plugins.hpp:1340-1342
// with the SAME decision sequence per query ("exact-order mode"): neighbours are taken in stored
// order, distances use strict '<' exactly where the reference does, a new element is placed before
// equal ones and the last one is evicted.  On tie-free data the returned ids are those of the
// reference on the same graph.  (The candidate queue is kept as "unexpanded" flags on the top list:
// an element evicted from `top` can never be popped again before the reference's stop test
// index.hpp:3445 fires, so the two formulations visit the same nodes; they differ only for exact
// distance ties at the eviction boundary.)
//
// Mapping to the machine: one CTA (4 warps) walks one query at a time; CTAs are persistent and pull
// queries from an atomic counter.  Per expanded node, warp 0 reads the 128-B adjacency line, filters
// it against the CTA's visited bitmap in HBM and compacts the unseen ids; then every warp streams
// its share of the neighbour rows HBM -> shared memory with 1-D bulk async copies (TMA engine,
// mbarrier completion), each warp refilling its own ring slots, reduces q.row with warp shuffles and
// warp 0 finally replays the reference's sequential insertion.  Rows are never re-read; the only
// HBM traffic besides rows is the adjacency line per pop and the bitmap words.
#include <cuda_runtime.h>
#include <float.h>

#include <type_traits>

#include "distance.cuh"
#include "engine.h"

namespace lb200 {

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kDone = 0xFFFFFFFFu;

struct Ctrl {
    uint32_t q;        // current query index
    uint32_t n;        // number of candidates in cand_id[], or kDone
    uint32_t ntouched; // entries in the touched list
    uint32_t overflow; // touched list overflowed -> full bitmap clear
};

struct Smem {
    uint8_t* ring;
    uint64_t* full;
    float* top_d;
    uint32_t* top_i;
    uint32_t* cand_id;
    float* cand_d;
    Ctrl* ctrl;
};

__host__ __device__ inline size_t smem_layout(uint32_t R, uint32_t row_bytes, uint32_t L, uint32_t M0, size_t* off_full,
                                              size_t* off_topd, size_t* off_topi, size_t* off_cid, size_t* off_cd,
                                              size_t* off_ctrl) {
    size_t o = (size_t)R * row_bytes;
    o = (o + 15) & ~(size_t)15;
    *off_full = o, o += (size_t)R * 8;
    *off_topd = o, o += (size_t)L * 4;
    *off_topi = o, o += (size_t)L * 4;
    *off_cid = o, o += (size_t)M0 * 4;
    *off_cd = o, o += (size_t)M0 * 4;
    o = (o + 15) & ~(size_t)15;
    *off_ctrl = o, o += sizeof(Ctrl);
    return o;
}

// Warp-cooperative sorted insert == sorted_buffer_gt::insert (index.hpp:752-763) for a full warp.
// Precondition (checked by the caller, index.hpp:3470): size < L || d < top_d[size-1].
__device__ __forceinline__ void top_insert(float* td, uint32_t* ti, uint32_t& size, uint32_t& cursor, uint32_t L, float d,
                                           uint32_t id, int lane) {
    uint32_t pos = 0;
    for (uint32_t b = 0; b < size; b += 32) {
        uint32_t e = b + lane;
        bool less = e < size && td[e] < d; // lower_bound: elements strictly below d form a prefix
        pos += __popc(__ballot_sync(0xffffffffu, less));
    }
    uint32_t last = (size == L) ? L - 1 : size; // index the old tail moves to (evicting when full)
    for (int hi = (int)last; hi > (int)pos; hi -= 32) {
        int idx = hi - lane;
        bool act = idx > (int)pos;
        float vd = 0.f;
        uint32_t vi = 0;
        if (act)
            vd = td[idx - 1], vi = ti[idx - 1];
        __syncwarp();
        if (act)
            td[idx] = vd, ti[idx] = vi;
        __syncwarp();
    }
    if (lane == 0)
        td[pos] = d, ti[pos] = id;
    __syncwarp();
    size = last + 1;
    if (pos <= cursor)
        cursor = pos;
}

template <int DM, int SK, int NQ> struct Walker {
    const SearchLaunch& p;
    Smem sm;
    uint4 qreg[NQ];
    float a2;
    uint32_t phase_bits;
    uint32_t nchunks, R, SPW;
    int warp, lane;

    __device__ __forceinline__ void issue(uint32_t slot, uint32_t id) {
        uint64_t* bar = &sm.full[slot];
        mbar_arrive_expect_tx(bar, p.g.row_bytes);
        bulk_g2s(sm.ring + (size_t)slot * p.g.row_bytes, p.g.vectors + (size_t)id * p.g.row_bytes, p.g.row_bytes, bar);
    }

    // distances from the query to cand_id[0..n) -> cand_d[0..n); candidate j is served by warp j%4,
    // which owns ring slots {warp, warp+4, ...} and refills each slot itself as soon as it has read it.
    __device__ __forceinline__ void eval(uint32_t n) {
        const uint32_t T = n > (uint32_t)warp ? (n - warp + kWarps - 1) / kWarps : 0;
        if ((uint32_t)lane < min(T, SPW))
            issue(warp + kWarps * lane, sm.cand_id[warp + kWarps * lane]);
        uint32_t si = 0;
        for (uint32_t t = 0; t < T; ++t) {
            const uint32_t slot = warp + kWarps * si;
            mbar_wait(&sm.full[slot], (phase_bits >> si) & 1u);
            phase_bits ^= 1u << si;
            const uint4* row = reinterpret_cast<const uint4*>(sm.ring + (size_t)slot * p.g.row_bytes);
            DistAcc<DM, SK> acc;
            acc.reset();
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                uint32_t c = lane + 32 * i;
                if (c < nchunks) {
                    uint4 r = row[c];
                    accum_chunk<DM, SK>(acc, qreg[i], r);
                }
            }
            float d = finish_distance<DM, SK>(acc, a2);
            if (lane == 0)
                sm.cand_d[warp + kWarps * t] = d;
            __syncwarp();
            if (t + SPW < T && lane == 0) {
                fence_proxy_async(); // order our generic-proxy reads of the slot before the async refill
                issue(slot, sm.cand_id[warp + kWarps * (t + SPW)]);
            }
            si = (si + 1 == SPW) ? 0 : si + 1;
        }
    }
};

template <int DM, int SK, int NQ>
__global__ void __launch_bounds__(kThreads) hnsw_search_kernel(const SearchLaunch p, const uint32_t R) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    size_t o_full, o_td, o_ti, o_cid, o_cd, o_ctrl;
    smem_layout(R, p.g.row_bytes, p.L, p.g.M0, &o_full, &o_td, &o_ti, &o_cid, &o_cd, &o_ctrl);

    Walker<DM, SK, NQ> w{p};
    w.sm.ring = smem_raw;
    w.sm.full = reinterpret_cast<uint64_t*>(smem_raw + o_full);
    w.sm.top_d = reinterpret_cast<float*>(smem_raw + o_td);
    w.sm.top_i = reinterpret_cast<uint32_t*>(smem_raw + o_ti);
    w.sm.cand_id = reinterpret_cast<uint32_t*>(smem_raw + o_cid);
    w.sm.cand_d = reinterpret_cast<float*>(smem_raw + o_cd);
    w.sm.ctrl = reinterpret_cast<Ctrl*>(smem_raw + o_ctrl);
    w.warp = threadIdx.x >> 5, w.lane = threadIdx.x & 31;
    w.nchunks = p.g.row_bytes / 16;
    w.R = R, w.SPW = R / kWarps;
    w.phase_bits = 0;
    w.a2 = 0.f;
    const int warp = w.warp, lane = w.lane;
    Smem& sm = w.sm;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < R; ++s)
            mbar_init(&sm.full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t* vis = p.s.visited + (size_t)blockIdx.x * p.s.words_per_cta;
    uint32_t* touched = p.s.touched + (size_t)blockIdx.x * p.s.touched_cap;
    const uint32_t L = p.L, M = p.g.M, M0 = p.g.M0;
    unsigned long long st_dist = 0, st_pops = 0, st_hops = 0; // meaningful in thread 0

    for (;;) {
        if (threadIdx.x == 0)
            sm.ctrl->q = (uint32_t)atomicAdd(&p.s.counters[0], 1ull);
        __syncthreads();
        const uint32_t q = sm.ctrl->q;
        if (q >= p.nq)
            break;

        { // query -> registers (every warp keeps its own copy)
            const uint4* qg = reinterpret_cast<const uint4*>(p.queries + (size_t)q * p.query_stride);
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                uint32_t c = lane + 32 * i;
                w.qreg[i] = c < w.nchunks ? __ldg(qg + c) : make_uint4(0, 0, 0, 0);
                part += query_norm_chunk<DM, SK>(w.qreg[i]);
            }
            if constexpr (DM == DM_COS)
                w.a2 = warp_sum(part);
        }

        // ---- entry point ----
        if (threadIdx.x == 0)
            sm.cand_id[0] = p.g.entry;
        __syncthreads();
        w.eval(1);
        __syncthreads();
        uint32_t cur = p.g.entry;
        float cur_d = sm.cand_d[0];
        st_dist += 1;

        // ---- greedy descent, levels max_level..1 (search_for_one_) ----
        for (int level = p.g.max_level; level >= 1; --level) {
            for (;;) {
                __syncthreads(); // everyone is done reading cand_d of the previous pass
                if (warp == 0) {
                    const uint32_t* list = p.g.upper_adj + ((size_t)__ldg(p.g.upper_ref + cur) + (level - 1)) * M;
                    uint32_t n = 0;
                    for (uint32_t off = 0; off < M; off += 32) {
                        uint32_t id = (off + lane < M) ? __ldg(list + off + lane) : kNoNeighbor;
                        bool valid = id != kNoNeighbor;
                        uint32_t m = __ballot_sync(0xffffffffu, valid);
                        if (valid)
                            sm.cand_id[n + __popc(m & ((1u << lane) - 1u))] = id;
                        n += __popc(m);
                    }
                    if (lane == 0)
                        sm.ctrl->n = n;
                }
                __syncthreads();
                const uint32_t n = sm.ctrl->n;
                w.eval(n);
                __syncthreads();
                // one pass of index.hpp:3304-3311: first strict improvement chain == first minimum below cur_d
                float best = cur_d;
                int bi = -1;
                for (uint32_t j = 0; j < n; ++j) {
                    float d = sm.cand_d[j];
                    if (d < best)
                        best = d, bi = (int)j;
                }
                st_dist += n, st_hops += 1;
                if (bi < 0)
                    break;
                cur = sm.cand_id[bi], cur_d = best;
            }
        }

        // ---- base layer beam (search_to_find_in_base_) ----
        uint32_t size = 0, cursor = 0, ntouched = 0; // warp-0 uniform
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) {
                sm.top_d[0] = cur_d, sm.top_i[0] = cur;
                atomicOr(&vis[cur >> 5], 1u << (cur & 31));
                touched[0] = cur >> 5;
                sm.ctrl->overflow = 0;
            }
            size = 1, cursor = 0, ntouched = 1;
            // the reference measures the start node a second time here (index.hpp:3436); the value is
            // already known, so only the counter (which defines the algorithmic work) is advanced
            st_dist += 1;
            __syncwarp();
        }
        for (;;) {
            __syncthreads(); // (A) insertion of the previous round is complete
            if (warp == 0) {
                if (cursor >= size) {
                    if (lane == 0)
                        sm.ctrl->n = kDone;
                } else {
                    const uint32_t c = sm.top_i[cursor];
                    __syncwarp();
                    if (lane == 0)
                        sm.top_i[cursor] = c | kExpandedBit;
                    __syncwarp();
                    { // advance cursor to the next unexpanded entry
                        uint32_t nxt = size;
                        for (uint32_t b = cursor + 1; b < size; b += 32) {
                            uint32_t e = b + lane;
                            bool un = e < size && !(sm.top_i[e] & kExpandedBit);
                            uint32_t m = __ballot_sync(0xffffffffu, un);
                            if (m) {
                                nxt = b + __ffs(m) - 1;
                                break;
                            }
                        }
                        cursor = nxt;
                    }
                    const uint32_t* list = p.g.adj0 + (size_t)c * M0;
                    uint32_t n = 0;
                    for (uint32_t off = 0; off < M0; off += 32) {
                        uint32_t id = (off + lane < M0) ? __ldg(list + off + lane) : kNoNeighbor;
                        bool valid = id != kNoNeighbor;
                        if (!__any_sync(0xffffffffu, valid))
                            break;
                        // duplicates inside one list are legal in reference graphs (refine_ padding,
                        // index.hpp:3554-3558): only the first occurrence may be "unseen"
                        uint32_t peers = __match_any_sync(0xffffffffu, id);
                        bool first = valid && ((uint32_t)(__ffs(peers) - 1) == (uint32_t)lane);
                        bool fresh = false;
                        if (first) {
                            uint32_t bit = 1u << (id & 31);
                            fresh = !(atomicOr(&vis[id >> 5], bit) & bit);
                        }
                        uint32_t m = __ballot_sync(0xffffffffu, fresh);
                        uint32_t rank = __popc(m & ((1u << lane) - 1u));
                        if (fresh) {
                            sm.cand_id[n + rank] = id;
                            if (ntouched + rank < p.s.touched_cap)
                                touched[ntouched + rank] = id >> 5;
                            prefetch_l2(p.g.adj0 + (size_t)id * M0); // its own adjacency line, for when it is popped
                        }
                        n += __popc(m);
                        ntouched += __popc(m);
                    }
                    if (lane == 0)
                        sm.ctrl->n = n;
                    st_pops += 1;
                }
            }
            __syncthreads(); // (B)
            const uint32_t n = sm.ctrl->n;
            if (n == kDone)
                break;
            w.eval(n);
            __syncthreads(); // (C)
            if (warp == 0) {
                st_dist += n;
                for (uint32_t j = 0; j < n; ++j) {
                    const float d = sm.cand_d[j];
                    // index.hpp:3470: top.size() < top_limit || successor_dist < radius
                    if (size < L || d < sm.top_d[size - 1])
                        top_insert(sm.top_d, sm.top_i, size, cursor, L, d, sm.cand_id[j], lane);
                }
            }
        }

        // ---- results: top is ascending; shrink(k); keys (index.hpp:2722-2723, 2426-2433) ----
        if (warp == 0) {
            const uint32_t found = min(size, p.k);
            for (uint32_t i = lane; i < p.k; i += 32) {
                uint64_t key = ~0ull;
                float d = INFINITY;
                if (i < found) {
                    key = __ldg(p.g.keys + (sm.top_i[i] & ~kExpandedBit));
                    d = sm.top_d[i];
                }
                p.out_keys[(size_t)q * p.k + i] = key;
                p.out_dists[(size_t)q * p.k + i] = d;
            }
            if (lane == 0) {
                if (p.out_counts)
                    p.out_counts[q] = found;
                sm.ctrl->ntouched = ntouched;
            }
        }
        __syncthreads();
        { // un-visit: only the words this query touched
            const uint32_t nt = sm.ctrl->ntouched;
            if (nt <= p.s.touched_cap) {
                for (uint32_t i = threadIdx.x; i < nt; i += kThreads)
                    vis[touched[i]] = 0u;
            } else {
                for (size_t i = threadIdx.x; i < p.s.words_per_cta; i += kThreads)
                    vis[i] = 0u;
            }
        }
        // the next iteration's first __syncthreads orders these stores before any new atomicOr
    }

    if (threadIdx.x == 0) {
        atomicAdd(&p.s.counters[1], st_dist);
        atomicAdd(&p.s.counters[2], st_pops);
        atomicAdd(&p.s.counters[3], st_hops);
    }
}

uint32_t pick_ring_slots(uint32_t row_bytes) {
    uint32_t r = (48u * 1024u / row_bytes) & ~3u;
    if (r < 4)
        r = 4;
    if (r > 32)
        r = 32;
    return r;
}

int pick_nq(uint32_t row_bytes) {
    uint32_t need = (row_bytes / 16 + 31) / 32;
    const int opts[] = {1, 2, 3, 4, 6, 8, 12, 16};
    for (int o : opts)
        if ((uint32_t)o >= need)
            return o;
    return -1;
}

template <int DM, int SK, int NQ> void launch_one(const SearchLaunch& p, uint32_t R, size_t smem, uint32_t grid, cudaStream_t stream) {
    auto kern = hnsw_search_kernel<DM, SK, NQ>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, stream>>>(p, R);
    LB_CUDA(cudaGetLastError());
    count_launch();
}

template <int DM, int SK, int NQ> int occupancy_one(size_t smem) {
    auto kern = hnsw_search_kernel<DM, SK, NQ>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = 0;
    LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, kThreads, smem));
    return blocks;
}

// dispatch over (DM, SK, NQ); `fn` is a generic lambda taking three integral_constants
template <typename Fn> void dispatch(int dm, int sk, int nq, Fn&& fn) {
#define LB_NQ_CASES(DMv, SKv)                                                                                          \
    switch (nq) {                                                                                                      \
    case 1: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 1>{}); return;   \
    case 2: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 2>{}); return;   \
    case 3: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 3>{}); return;   \
    case 4: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 4>{}); return;   \
    case 6: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 6>{}); return;   \
    case 8: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 8>{}); return;   \
    case 12: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 12>{}); return; \
    case 16: fn(std::integral_constant<int, DMv>{}, std::integral_constant<int, SKv>{}, std::integral_constant<int, 16>{}); return; \
    default: break;                                                                                                    \
    }
    if (dm == DM_L2SQ && sk == SK_F32) {
        LB_NQ_CASES(DM_L2SQ, SK_F32)
    } else if (dm == DM_COS && sk == SK_F32) {
        LB_NQ_CASES(DM_COS, SK_F32)
    } else if (dm == DM_L2SQ && sk == SK_F16) {
        LB_NQ_CASES(DM_L2SQ, SK_F16)
    } else if (dm == DM_COS && sk == SK_F16) {
        LB_NQ_CASES(DM_COS, SK_F16)
    } else if (dm == DM_L2SQ && sk == SK_I8) {
        LB_NQ_CASES(DM_L2SQ, SK_I8)
    } else if (dm == DM_COS && sk == SK_I8) {
        LB_NQ_CASES(DM_COS, SK_I8)
    } else if (dm == DM_HAMMING && sk == SK_B1) {
        LB_NQ_CASES(DM_HAMMING, SK_B1)
    }
#undef LB_NQ_CASES
    throw CudaError("search: unsupported metric / scalar kind / dimensionality combination");
}

size_t search_smem(uint32_t R, uint32_t row_bytes, uint32_t L, uint32_t M0) {
    size_t a, b, c, d, e, f;
    return smem_layout(R, row_bytes, L, M0, &a, &b, &c, &d, &e, &f);
}

} // namespace

uint32_t search_max_ctas(int dist_mode, int scalar_kind, uint32_t row_bytes, uint32_t L, uint32_t M0, bool pq) {
    (void)pq;
    const uint32_t R = pick_ring_slots(row_bytes);
    const int nq = pick_nq(row_bytes);
    if (nq < 0)
        throw CudaError("search: vectors wider than 8192 bytes are not supported");
    const size_t smem = search_smem(R, row_bytes, L, M0);
    int occ = 0;
    dispatch(dist_mode, scalar_kind, nq, [&](auto dm, auto sk, auto n) { occ = occupancy_one<decltype(dm)::value, decltype(sk)::value, decltype(n)::value>(smem); });
    if (occ < 1)
        throw CudaError("search: kernel does not fit on an SM (ef/k too large for shared memory?)");
    return (uint32_t)occ * (uint32_t)device_sm_count();
}

void launch_search(int dist_mode, int scalar_kind, const SearchLaunch& p, cudaStream_t stream) {
    const uint32_t R = pick_ring_slots(p.g.row_bytes);
    const int nq = pick_nq(p.g.row_bytes);
    const size_t smem = search_smem(R, p.g.row_bytes, p.L, p.g.M0);
    const uint32_t grid = p.s.ctas < p.nq ? p.s.ctas : p.nq;
    dispatch(dist_mode, scalar_kind, nq,
             [&](auto dm, auto sk, auto n) { launch_one<decltype(dm)::value, decltype(sk)::value, decltype(n)::value>(p, R, smem, grid, stream); });
}

} // namespace lb200
