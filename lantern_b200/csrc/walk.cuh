// lantern_b200 -- the graph walker shared by the search and build kernels.
//
// One CTA (4 warps) executes, for one "value" (a query or a vector being inserted), the reference's
//   search_for_one_            U/include/usearch/index.hpp:3277-3316   -> Walker::greedy
//   search_to_find_in_base_    :3400-3485                              -> Walker::beam(level 0)
//   search_to_insert_          :3324-3392                              -> Walker::beam(level l)
//   refine_                    :3515-3561                              -> Walker::refine
// with the same decision sequence (stored neighbour order, strict '<', insert-before-equal, evict-last).
// The candidate queue (`next`, a max-heap on -distance in the reference) is represented by "unexpanded"
// flags on the sorted `top` list: an element evicted from `top` has distance >= radius and can never be
// popped before the stop test (:3445 / :3351) fires, so both formulations expand the same nodes; they can
// differ only for exact distance ties.
//
// Data movement: neighbour rows travel HBM -> shared memory as 1-D bulk async copies (TMA engine,
// `cp.async.bulk`, mbarrier complete_tx).  Candidate j of a batch is served by warp j%4, which owns the
// ring slots {warp, warp+4, ...} and refills a slot itself right after reading it, so a warp's pipeline
// needs no cross-warp synchronisation.  Distances are reduced with warp shuffles (distance.cuh).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "distance.cuh"
#include "engine.h"

namespace lb200 {

constexpr int kWalkThreads = 128;
constexpr int kWalkWarps = kWalkThreads / 32;
constexpr uint32_t kDone = 0xFFFFFFFFu;
constexpr uint32_t kIdMask = 0x7FFFFFFFu;

struct WalkCtrl {
    uint32_t item;     // current work item (query / node / segment)
    uint32_t n;        // candidates in cand_id[], or kDone
    uint32_t ntouched; // entries in the touched list
    uint32_t top_size; // size of the top list after a beam
};

struct WalkSmem {
    uint8_t* ring;
    uint64_t* full;
    float* top_d;
    uint32_t* top_i;
    uint32_t* cand_id;
    float* cand_d;
    uint32_t* limbo; // unexpanded entries evicted from the top list at exactly the current radius (distance ties)
    WalkCtrl* ctrl;
};

struct WalkLayout {
    size_t full, top_d, top_i, cand_id, cand_d, limbo, ctrl, total;
};
constexpr uint32_t kLimboCap = 64;

// `stage_bytes` of value/row staging, `nbar` mbarriers, a top list of `top_cap` entries, candidate arrays of `cand_cap`.
__host__ __device__ inline WalkLayout walk_layout_bytes(size_t stage_bytes, uint32_t nbar, uint32_t top_cap, uint32_t cand_cap) {
    WalkLayout l;
    size_t o = stage_bytes;
    o = (o + 15) & ~(size_t)15;
    l.full = o, o += (size_t)nbar * 8;
    l.top_d = o, o += (size_t)top_cap * 4;
    l.top_i = o, o += (size_t)top_cap * 4;
    l.cand_id = o, o += (size_t)cand_cap * 4;
    l.cand_d = o, o += (size_t)cand_cap * 4;
    l.limbo = o, o += (size_t)kLimboCap * 4;
    o = (o + 15) & ~(size_t)15;
    l.ctrl = o, o += sizeof(WalkCtrl);
    l.total = o;
    return l;
}
// R ring slots of row_bytes (plain rows)
__host__ __device__ inline WalkLayout walk_layout(uint32_t R, uint32_t row_bytes, uint32_t top_cap, uint32_t cand_cap) {
    return walk_layout_bytes((size_t)R * row_bytes, R, top_cap, cand_cap);
}
// PQ: look-up table [nsub][ncent] (+ 4 floats: the value's squared norm travels with a precomputed table) + value staging
// [dims] + per-warp partial sums [warps][2][64]; one mbarrier (bulk copy of a precomputed table)
constexpr uint32_t kPqPartPerWarp = 64;
__host__ __device__ inline size_t pq_table_floats(uint32_t nsub, uint32_t lut_width) { return (size_t)nsub * lut_width + 4; }
// `value_floats` = dims when the kernel builds tables itself from raw vectors (build; search without precomputed tables), else 0
__host__ __device__ inline WalkLayout walk_layout_pq(uint32_t nsub, uint32_t lut_width, uint32_t value_floats, uint32_t top_cap, uint32_t cand_cap) {
    return walk_layout_bytes((pq_table_floats(nsub, lut_width) + value_floats + (size_t)kWalkWarps * 2 * kPqPartPerWarp) * 4, 1, top_cap, cand_cap);
}
__host__ __device__ inline uint32_t pq_value_floats(const GraphView& g) { return g.pq_query_tables ? 0u : g.dims; }

inline uint32_t pick_ring_slots(uint32_t row_bytes) {
    uint32_t budget = 24u * 1024u; // 8 slots of a d=768 f32 row: 8 CTAs/SM; measured best on B200 (profiles/)
    if (const char* e = getenv("LB200_RING_BYTES")) // tuning knob (bytes of row staging per CTA)
        budget = (uint32_t)atoi(e);
    uint32_t r = (budget / row_bytes) & ~3u;
    if (r < 4)
        r = 4;
    if (r > 16) // narrow rows: more resident CTAs beat deeper rings (768-byte rows: 16 slots +15 % over 32, profiles/)
        r = 16;
    return r;
}

inline int pick_nq(uint32_t row_bytes) {
    const uint32_t need = (row_bytes / 16 + 31) / 32;
    const int opts[] = {1, 2, 3, 4, 6, 8, 12, 16};
    for (int o : opts)
        if ((uint32_t)o >= need)
            return o;
    return -1;
}

// Warp-cooperative sorted insert == sorted_buffer_gt::insert (index.hpp:752-763): position = lower_bound
// (new element goes BEFORE equal ones), tail evicted when the list is full.
// Precondition (index.hpp:3470): size < L || d < top_d[size-1].
// `ev_d`/`ev_i` receive the evicted tail (ev_i = kNoNeighbor when nothing was evicted).
__device__ __forceinline__ void top_insert(float* td, uint32_t* ti, uint32_t& size, uint32_t& cursor, uint32_t L, float d,
                                           uint32_t id, int lane, float& ev_d, uint32_t& ev_i) {
    ev_i = kNoNeighbor, ev_d = 0.f;
    if (size == L)
        ev_d = td[L - 1], ev_i = ti[L - 1];
    __syncwarp();
    uint32_t pos = 0;
    for (uint32_t b = 0; b < size; b += 32) {
        uint32_t e = b + lane;
        bool less = e < size && td[e] < d;
        pos += __popc(__ballot_sync(0xffffffffu, less));
    }
    const uint32_t last = (size == L) ? L - 1 : size;
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
    __syncwarp(); // the position scan's reads are ordered before this write even when nothing was shifted (racecheck)
    if (lane == 0)
        td[pos] = d, ti[pos] = id;
    __syncwarp();
    size = last + 1;
    if (pos <= cursor)
        cursor = pos;
}
__device__ __forceinline__ void top_insert(float* td, uint32_t* ti, uint32_t& size, uint32_t& cursor, uint32_t L, float d,
                                           uint32_t id, int lane) {
    float ev_d;
    uint32_t ev_i;
    top_insert(td, ti, size, cursor, L, d, id, lane, ev_d, ev_i);
}

// ---- the top-`L` list of one walk (warp 0 only) ------------------------------------------------------------------
// sorted_buffer_gt (index.hpp:668-780): ascending by distance, insert at lower_bound (before equal elements), evict the
// tail when full; entries carry an "expanded" flag (kExpandedBit) that replaces the reference's separate candidate heap.
struct TopSmem { // arrays in shared memory
    float* td;
    uint32_t* ti;
    uint32_t size, cursor, L;
    __device__ __forceinline__ void init(const WalkSmem& sm, uint32_t cap, float d, uint32_t id, int lane) {
        td = sm.top_d, ti = sm.top_i, L = cap, size = 1, cursor = 0;
        if (lane == 0)
            td[0] = d, ti[0] = id;
        __syncwarp();
    }
    // closest unexpanded entry -> c (marked expanded); false when none is left
    __device__ __forceinline__ bool pop(uint32_t& c, int lane) {
        if (cursor >= size)
            return false;
        c = ti[cursor];
        __syncwarp();
        if (lane == 0)
            ti[cursor] = c | kExpandedBit;
        __syncwarp();
        uint32_t nxt = size;
        for (uint32_t b = cursor + 1; b < size; b += 32) {
            uint32_t e = b + lane;
            bool un = e < size && !(ti[e] & kExpandedBit);
            uint32_t m = __ballot_sync(0xffffffffu, un);
            if (m) {
                nxt = b + __ffs(m) - 1;
                break;
            }
        }
        cursor = nxt;
        return true;
    }
    __device__ __forceinline__ float radius(int) const { return td[size - 1]; }
    __device__ __forceinline__ void insert(float d, uint32_t id, int lane, float& ev_d, uint32_t& ev_i) {
        top_insert(td, ti, size, cursor, L, d, id, lane, ev_d, ev_i);
    }
    __device__ __forceinline__ void flush(int) {}
};

// state shared by every evaluator / walker
struct WalkBase {
    const GraphView& g; // the kernel's own (__grid_constant__) parameter: fields are read from the constant bank
    WalkSmem sm;
    uint32_t* vis;
    uint32_t* touched;
    uint32_t touched_cap;
    size_t words_per_cta;
    int warp, lane;
    uint32_t st_dist, st_pops, st_hops; // per-launch work counters; thread 0's copy is the one that is reported
    uint32_t st_limbo_drop;             // equal-distance candidates that did not fit in limbo (the walk then expands fewer nodes)

    __device__ __forceinline__ explicit WalkBase(const GraphView& gv) : g(gv) {}
    __device__ __forceinline__ void init_base(uint8_t* smem_raw, const WalkLayout& lay, const SearchScratch& s) {
        sm.ring = smem_raw;
        sm.full = reinterpret_cast<uint64_t*>(smem_raw + lay.full);
        sm.top_d = reinterpret_cast<float*>(smem_raw + lay.top_d);
        sm.top_i = reinterpret_cast<uint32_t*>(smem_raw + lay.top_i);
        sm.cand_id = reinterpret_cast<uint32_t*>(smem_raw + lay.cand_id);
        sm.cand_d = reinterpret_cast<float*>(smem_raw + lay.cand_d);
        sm.limbo = reinterpret_cast<uint32_t*>(smem_raw + lay.limbo);
        sm.ctrl = reinterpret_cast<WalkCtrl*>(smem_raw + lay.ctrl);
        warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        st_dist = st_pops = st_hops = 0;
        st_limbo_drop = 0;
        vis = s.visited + (size_t)blockIdx.x * s.words_per_cta;
        touched = s.touched + (size_t)blockIdx.x * s.touched_cap;
        touched_cap = s.touched_cap;
        words_per_cta = s.words_per_cta;
    }
};

// ---- evaluator for plain rows (f32 / f16 / i8 / b1): value in registers, rows through the bulk-copy ring --------
template <int DM, int SK, int NQ> struct RowEval : WalkBase {
    static __host__ __device__ WalkLayout layout(const GraphView& g, uint32_t R, uint32_t top_cap, uint32_t cand_cap) {
        return walk_layout(R, g.row_bytes, top_cap, cand_cap);
    }
    uint4 qreg[NQ];
    float a2;
    uint32_t phase_bits;
    uint32_t nchunks, R, SPW, NW;

    __device__ __forceinline__ explicit RowEval(const GraphView& gv) : WalkBase(gv) {}
    __device__ __forceinline__ void init(uint8_t* smem_raw, const WalkLayout& lay, uint32_t ring_slots, const SearchScratch& s) {
        init_base(smem_raw, lay, s);
        nchunks = g.row_bytes / 16;
        NW = blockDim.x >> 5; // 4 warps per query, or 2 for narrow rows (search.cu): uniform across the CTA
        R = ring_slots, SPW = ring_slots / NW;
        phase_bits = 0;
        a2 = 0.f;
        if (threadIdx.x == 0) {
            for (uint32_t i = 0; i < R; ++i)
                mbar_init(&sm.full[i], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }

    // value -> registers; every warp keeps its own copy.  `row` points to a 16-byte padded row in global memory.
    __device__ __forceinline__ void load_value(const uint8_t* row) {
        const uint4* qg = reinterpret_cast<const uint4*>(row);
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            uint32_t c = lane + 32 * i;
            qreg[i] = c < nchunks ? __ldg(qg + c) : make_uint4(0, 0, 0, 0);
            part = norm_add(part, query_norm_chunk<DM, SK>(qreg[i]));
        }
        if constexpr (DM == DM_COS)
            a2 = warp_sum(part);
    }
    // a stored node becomes the value (refine_: candidate vs accepted; reverse links: target vs its neighbours)
    __device__ __forceinline__ void load_node(uint32_t id) { load_value(g.vectors + (size_t)id * g.row_bytes); }
    __device__ __forceinline__ void load_query(uint32_t, const uint8_t* row) { load_value(row); }

    __device__ __forceinline__ void issue(uint32_t slot, uint32_t id) {
        uint64_t* bar = &sm.full[slot];
        mbar_arrive_expect_tx(bar, g.row_bytes);
        bulk_g2s(sm.ring + (size_t)slot * g.row_bytes, g.vectors + (size_t)id * g.row_bytes, g.row_bytes, bar);
    }

    __device__ __forceinline__ float row_distance(const uint4* row) const {
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
        return finish_distance<DM, SK>(acc, a2);
    }

    // distances value -> cand_id[0..n) into cand_d[0..n).  Callers bracket it with __syncthreads().
    __device__ __forceinline__ void eval(uint32_t n) {
        const uint32_t T = n > (uint32_t)warp ? (n - warp + NW - 1) / NW : 0;
        if ((uint32_t)lane < min(T, SPW))
            issue(warp + NW * lane, sm.cand_id[warp + NW * lane]);
        uint32_t si = 0;
        for (uint32_t t = 0; t < T; ++t) {
            const uint32_t slot = warp + NW * si;
            mbar_wait(&sm.full[slot], (phase_bits >> si) & 1u);
            phase_bits ^= 1u << si;
            const float d = row_distance(reinterpret_cast<const uint4*>(sm.ring + (size_t)slot * g.row_bytes));
            if (lane == 0)
                sm.cand_d[warp + NW * t] = d;
            __syncwarp();
            if (t + SPW < T && lane == 0) {
                fence_proxy_async(); // our generic-proxy reads of the slot precede the async refill
                issue(slot, sm.cand_id[warp + NW * (t + SPW)]);
            }
            si = (si + 1 == SPW) ? 0 : si + 1;
        }
    }
};

// ---- evaluator for PQ-coded rows: asymmetric distance through a per-value look-up table ----------------------
// Reference semantics (lantern_storage.hpp:137-149,264-267; index_dense.hpp:341-358): every stored operand is
// decompressed (concatenation of centroid slices), the query / new vector never is.  Because a decoded vector is a
// concatenation of centroid slices, d(value, decode(code)) = sum_s lut[s][code_s] with
//   l2sq: lut[s][c] = |value_s - centroid_{c,s}|^2          cos: lut[s][c] = value_s . centroid_{c,s}  (+ norm tables)
// which reproduces the reference up to fp32 summation order (verified 2e-7 relative).  When the value
// is itself a stored node (refine_, reverse links) its table rows are copied from the centroid-pair table `pq_pair`
// [nsub][ncent][ncent] precomputed at index creation.  Code rows (nsub bytes) are read straight from HBM/L2.
template <int DM> struct PqEval : WalkBase {
    static __host__ __device__ WalkLayout layout(const GraphView& g, uint32_t, uint32_t top_cap, uint32_t cand_cap) {
        return walk_layout_pq(g.num_subvectors, g.pq_lut_width, pq_value_floats(g), top_cap, cand_cap);
    }
    float* lut;  // shared [nsub][ncent] (+ 4)
    float* qbuf; // shared [dims]
    float* part; // shared [warps][2][64]: per-chunk partial sums of the candidates a warp is measuring
    float a2;
    uint32_t tphase;

    __device__ __forceinline__ explicit PqEval(const GraphView& gv) : WalkBase(gv) {}
    __device__ __forceinline__ void init(uint8_t* smem_raw, const WalkLayout& lay, uint32_t, const SearchScratch& s) {
        init_base(smem_raw, lay, s);
        lut = reinterpret_cast<float*>(smem_raw);
        qbuf = lut + pq_table_floats(g.num_subvectors, g.pq_lut_width);
        part = qbuf + pq_value_floats(g);
        a2 = 0.f;
        tphase = 0;
        if (threadIdx.x == 0) {
            mbar_init(&sm.full[0], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }

    // one table entry: |v_s - c|^2 (l2sq) or v_s . c (cos), dimensions in ascending order (the order pq_query_tables_kernel uses)
    __device__ __forceinline__ static float entry(const float* qs, const float* cen, uint32_t sd) {
        float acc = 0.f;
        if ((sd & 3u) == 0u) {
            for (uint32_t i = 0; i < sd; i += 4) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cen + i));
                if constexpr (DM == DM_COS) {
                    acc = fmaf(qs[i], c4.x, acc), acc = fmaf(qs[i + 1], c4.y, acc);
                    acc = fmaf(qs[i + 2], c4.z, acc), acc = fmaf(qs[i + 3], c4.w, acc);
                } else {
                    float d = qs[i] - c4.x;
                    acc = fmaf(d, d, acc);
                    d = qs[i + 1] - c4.y, acc = fmaf(d, d, acc);
                    d = qs[i + 2] - c4.z, acc = fmaf(d, d, acc);
                    d = qs[i + 3] - c4.w, acc = fmaf(d, d, acc);
                }
            }
        } else {
            for (uint32_t i = 0; i < sd; ++i) {
                const float cv = __ldg(cen + i);
                if constexpr (DM == DM_COS)
                    acc = fmaf(qs[i], cv, acc);
                else {
                    const float d = qs[i] - cv;
                    acc = fmaf(d, d, acc);
                }
            }
        }
        return acc;
    }

    // `row` = raw f32 vector (dims floats) in global memory
    __device__ __forceinline__ void load_value(const uint8_t* row) {
        const float* q = reinterpret_cast<const float*>(row);
        const uint32_t dims = g.dims, ncent = g.pq_lut_width, nsub = g.num_subvectors, sd = dims / nsub;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < dims; i += kWalkThreads)
            qbuf[i] = __ldg(q + i);
        __syncthreads();
        if constexpr (DM == DM_COS) {
            float p2 = 0.f;
            for (uint32_t i = lane; i < dims; i += 32)
                p2 += qbuf[i] * qbuf[i];
            a2 = warp_sum(p2);
        }
        for (uint32_t e = threadIdx.x; e < nsub * ncent; e += kWalkThreads) {
            const uint32_t s = e / ncent, c = e - s * ncent;
            lut[e] = entry(qbuf + (size_t)s * sd, g.codebook + (size_t)c * dims + (size_t)s * sd, sd);
        }
        __syncthreads();
    }

    // search: query `qi` of the batch.  When the batch's tables were precomputed (pq_query_tables_kernel, one dense launch at
    // full occupancy instead of nsub*ncent*subdim flops per query inside this 4-CTA/SM kernel) the table is one bulk copy away.
    __device__ __forceinline__ void load_query(uint32_t qi, const uint8_t* row) {
        if (!g.pq_query_tables) {
            load_value(row);
            return;
        }
        const uint32_t floats = (uint32_t)pq_table_floats(g.num_subvectors, g.pq_lut_width);
        __syncthreads(); // every look-up of the previous query is done
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&sm.full[0], floats * 4);
            bulk_g2s(lut, g.pq_query_tables + (size_t)qi * floats, floats * 4, &sm.full[0]);
        }
        mbar_wait(&sm.full[0], tphase);
        tphase ^= 1u;
        a2 = lut[(size_t)g.num_subvectors * g.pq_lut_width]; // |query|^2 (cos)
    }

    __device__ __forceinline__ void load_node(uint32_t id) {
        const uint32_t ncent = g.pq_lut_width, full = g.num_centroids, nsub = g.num_subvectors;
        const uint8_t* codes = g.vectors + (size_t)id * g.row_bytes;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < nsub * ncent; e += kWalkThreads) {
            const uint32_t s = e / ncent, c = e - s * ncent;
            lut[e] = __ldg(g.pq_pair + ((size_t)s * full + __ldg(codes + s)) * full + c);
        }
        if constexpr (DM == DM_COS) {
            float p2 = 0.f;
            for (uint32_t s = lane; s < nsub; s += 32)
                p2 += __ldg(g.pq_norm + (size_t)s * full + __ldg(codes + s));
            a2 = warp_sum(p2);
        }
        __syncthreads();
    }

    // 16 codes of one chunk -> partial sums
    __device__ __forceinline__ void chunk_sums(const uint4& v, uint32_t ch, float& acc, float& b2) const {
        const uint32_t ncent = g.pq_lut_width, full = g.num_centroids, nsub = g.num_subvectors;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        acc = 0.f, b2 = 0.f;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t s = ch * 16 + wi * 4 + b;
                if (s < nsub) {
                    const uint32_t c = (w[wi] >> (8 * b)) & 255u;
                    acc += lut[s * ncent + c];
                    if constexpr (DM == DM_COS)
                        b2 += __ldg(g.pq_norm + (size_t)s * full + c);
                }
            }
    }

    // distances value -> cand_id[0..n).  Candidate j belongs to warp j % 4.  A warp fetches the code rows of ALL its candidates of
    // a pass at once (16-byte chunks, two per lane, every load issued before the first look-up -- one HBM latency per round
    // instead of one per candidate), each lane sums the 16 look-ups of its chunks, and the chunk sums of a candidate are added
    // in ascending chunk order (a fixed order: the distance does not depend on which lane or pass served the candidate).
    __device__ __forceinline__ void eval(uint32_t n) {
        const uint32_t cpr = g.row_bytes / 16; // chunks per code row (6 at 96 subvectors)
        const uint32_t T = n > (uint32_t)warp ? (n - warp + kWalkWarps - 1) / kWalkWarps : 0;
        float* pa = part + (size_t)warp * 2 * kPqPartPerWarp;
        float* pb = pa + kPqPartPerWarp;
        if (cpr <= kPqPartPerWarp) {
            const uint32_t per_pass = kPqPartPerWarp / cpr; // whole candidates per pass
            for (uint32_t t0 = 0; t0 < T; t0 += per_pass) {
                const uint32_t cnt = min(per_pass, T - t0), chunks = cnt * cpr;
                uint4 v[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t idx = lane + 32 * u;
                    if (idx < chunks) {
                        const uint32_t t = idx / cpr, ch = idx - t * cpr;
                        const uint32_t id = sm.cand_id[warp + kWalkWarps * (t0 + t)];
                        v[u] = __ldg(reinterpret_cast<const uint4*>(g.vectors + (size_t)id * g.row_bytes) + ch);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t idx = lane + 32 * u;
                    if (idx < chunks) {
                        const uint32_t t = idx / cpr, ch = idx - t * cpr;
                        float acc, b2;
                        chunk_sums(v[u], ch, acc, b2);
                        pa[idx] = acc;
                        if constexpr (DM == DM_COS)
                            pb[idx] = b2;
                    }
                }
                __syncwarp();
                if ((uint32_t)lane < cnt) {
                    float acc = 0.f, b2 = 0.f;
                    for (uint32_t ch = 0; ch < cpr; ++ch) {
                        acc += pa[lane * cpr + ch];
                        if constexpr (DM == DM_COS)
                            b2 += pb[lane * cpr + ch];
                    }
                    float d = acc;
                    if constexpr (DM == DM_COS)
                        d = cos_from_parts(acc, a2, b2);
                    sm.cand_d[warp + kWalkWarps * (t0 + lane)] = d;
                }
                __syncwarp();
            }
        } else { // very wide code rows (> 1024 subvectors): one candidate at a time, chunks strided over the lanes
            for (uint32_t t = 0; t < T; ++t) {
                const uint32_t id = sm.cand_id[warp + kWalkWarps * t];
                float acc = 0.f, b2 = 0.f;
                for (uint32_t ch = lane; ch < cpr; ch += 32) {
                    float a1, b1;
                    chunk_sums(__ldg(reinterpret_cast<const uint4*>(g.vectors + (size_t)id * g.row_bytes) + ch), ch, a1, b1);
                    acc += a1, b2 += b1;
                }
                acc = warp_sum(acc);
                float d = acc;
                if constexpr (DM == DM_COS) {
                    b2 = warp_sum(b2);
                    d = cos_from_parts(acc, a2, b2);
                }
                if (lane == 0)
                    sm.cand_d[warp + kWalkWarps * t] = d;
            }
        }
    }
};

template <class E> struct WalkerT : E {
    __device__ __forceinline__ explicit WalkerT(const GraphView& gv) : E(gv) {}
    using E::eval;
    using E::g;
    using E::lane;
    using E::load_node;
    using E::sm;
    using E::st_dist;
    using E::st_hops;
    using E::st_pops;
    using E::st_limbo_drop;
    using E::touched;
    using E::touched_cap;
    using E::vis;
    using E::warp;
    using E::words_per_cta;

    __device__ __forceinline__ const uint32_t* list_of(uint32_t node, int level, uint32_t& width) const {
        if (level == 0) {
            width = g.M0;
            return g.adj0 + (size_t)node * g.M0;
        }
        width = g.M;
        return g.upper_adj + ((size_t)__ldg(g.upper_ref + node) + (level - 1)) * g.M;
    }

    // distance value -> one node (entry point)
    __device__ __forceinline__ float measure_one(uint32_t id) {
        __syncthreads();
        if (threadIdx.x == 0)
            sm.cand_id[0] = id;
        __syncthreads();
        eval(1);
        __syncthreads();
        st_dist += 1;
        return sm.cand_d[0];
    }

    // search_for_one_: levels from_level, from_level-1, ..., stop_level+1
    __device__ __forceinline__ void greedy(uint32_t& cur, float& cur_d, int from_level, int stop_level) {
        for (int level = from_level; level > stop_level; --level) {
            for (;;) {
                __syncthreads(); // everyone is done with cand_* of the previous pass
                if (warp == 0) {
                    uint32_t width;
                    const uint32_t* list = list_of(cur, level, width);
                    uint32_t n = 0;
                    for (uint32_t off = 0; off < width; off += 32) {
                        uint32_t id = (off + lane < width) ? __ldg(list + off + lane) : kNoNeighbor;
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
                eval(n);
                __syncthreads();
                // one pass of index.hpp:3304-3311: chain of strict improvements == first minimum below cur_d
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
    }

    // Beam search on one level with top list capacity L, starting from `start` (distance start_d already known:
    // the reference re-measures it, so the counter advances).  `skip` = node whose expansion is skipped
    // (index.hpp:3357 new_slot; kNoNeighbor for plain search).  Leaves the ascending top list in shared memory;
    // returns its size (uniform across the CTA).  Visited bits are cleared before returning.
    // (A register-resident top list for L <= 64 was measured on the B200: 72 registers/thread cost one resident CTA per SM
    //  and lost 2.5 % at batch 1024 -- the list stays in shared memory.)
    // `expand` = candidates expanded per round: 1 reproduces the reference's order exactly; p > 1 ("relaxed order") expands
    // the p closest unexpanded candidates together -- fewer serial rounds, a few per cent more distance evaluations, same
    // result at ef >= N, recall within the +-0.5 % window at finite ef.
    __device__ __forceinline__ uint32_t beam(int level, uint32_t start, float start_d, uint32_t L, uint32_t skip, uint32_t expand = 1) {
        return beam_impl<TopSmem>(level, start, start_d, L, skip, expand);
    }

    template <class Top>
    __device__ __forceinline__ uint32_t beam_impl(int level, uint32_t start, float start_d, uint32_t L, uint32_t skip, uint32_t expand) {
        uint32_t ntouched = 0; // warp-0 uniform
        Top top;
        // Distance ties at the eviction boundary: the reference's queue keeps an element after `top` evicted it, and still
        // expands it while its distance EQUALS the radius (the stop test index.hpp:3445 is a strict '>').  Such elements
        // wait in `limbo`; they all share one distance (the radius at the time) and die as soon as the radius shrinks.
        uint32_t limbo_n = 0;
        float limbo_d = 0.f;
        __syncthreads();
        if (warp == 0) {
            top.init(sm, L, start_d, start, lane);
            if (lane == 0) {
                atomicOr(&vis[start >> 5], 1u << (start & 31));
                touched[0] = start >> 5;
            }
            ntouched = 1;
            st_dist += 1; // index.hpp:3436 / :3343
            __syncwarp();
        }
        for (;;) {
            __syncthreads(); // (A) insertions of the previous round are complete
            if (warp == 0) {
                uint32_t n = 0, popped = 0;
                while (popped < expand) {
                    uint32_t c = kNoNeighbor;
                    if (!top.pop(c, lane) && limbo_n)
                        c = sm.limbo[--limbo_n]; // distance == radius: not beyond it, so the reference expands it too
                    if (c == kNoNeighbor)
                        break;
                    ++popped;
                    if (c == skip)
                        continue;
                    uint32_t width;
                    const uint32_t* list = list_of(c, level, width);
                    for (uint32_t off = 0; off < width; off += 32) {
                        uint32_t id = (off + lane < width) ? __ldg(list + off + lane) : kNoNeighbor;
                        bool valid = id != kNoNeighbor;
                        if (!__any_sync(0xffffffffu, valid))
                            break;
                        // duplicate ids inside one list are legal in reference graphs (refine_ padding,
                        // index.hpp:3554-3558): only the first occurrence can be "unseen"
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
                            if (ntouched + rank < touched_cap)
                                touched[ntouched + rank] = id >> 5;
                            if (level == 0 && (g.flags & 1u))
                                prefetch_l2(g.adj0 + (size_t)id * g.M0); // its adjacency line, for when it is popped
                        }
                        n += __popc(m);
                        ntouched += __popc(m);
                    }
                    st_pops += 1;
                }
                if (lane == 0)
                    sm.ctrl->n = popped ? n : kDone;
            }
            __syncthreads(); // (B)
            const uint32_t n = sm.ctrl->n;
            if (n == kDone)
                break;
            eval(n);
            __syncthreads(); // (C)
            if (warp == 0) {
                st_dist += n;
                // index.hpp:3470 / :3382: accepted iff top.size() < top_limit || successor_dist < radius.  The radius only
                // shrinks once the list is full, so a candidate that fails against the CURRENT radius can never pass
                // later in this round: filter those out in parallel, then replay the survivors in stored order.
                for (uint32_t base = 0; base < n; base += 32) {
                    const uint32_t j = base + lane;
                    const float dj = j < n ? sm.cand_d[j] : INFINITY;
                    float radius = top.radius(lane);
                    uint32_t m = __ballot_sync(0xffffffffu, j < n && (top.size < L || dj < radius));
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const float d = __shfl_sync(0xffffffffu, dj, b);
                        if (top.size < L || d < radius) {
                            const uint32_t id = sm.cand_id[base + b];
                            float ev_d;
                            uint32_t ev_i;
                            top.insert(d, id, lane, ev_d, ev_i);
                            if (level == 0 && (g.flags & 2u) && lane == 0)
                                prefetch_l2(g.adj0 + (size_t)id * g.M0);
                            radius = top.radius(lane);
                            if (limbo_n && radius < limbo_d)
                                limbo_n = 0; // the radius shrank below the waiting ties: they can never be expanded
                            if (ev_i != kNoNeighbor && !(ev_i & kExpandedBit) && ev_d == radius) {
                                if (limbo_n < kLimboCap) {
                                    if (lane == 0)
                                        sm.limbo[limbo_n] = ev_i;
                                    limbo_n++, limbo_d = radius;
                                    __syncwarp();
                                } else {
                                    this->st_limbo_drop += 1; // counted and reported (lb200_search_stats_t::limbo_overflows)
                                }
                            }
                        }
                    }
                }
            }
        }
        if (warp == 0) {
            top.flush(lane);
            if (lane == 0)
                sm.ctrl->ntouched = ntouched, sm.ctrl->top_size = top.size;
        }
        __syncthreads();
        { // un-visit only the words this walk touched
            const uint32_t nt = sm.ctrl->ntouched;
            if (nt <= touched_cap) {
                for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x)
                    vis[touched[i]] = 0u;
            } else {
                for (size_t i = threadIdx.x; i < words_per_cta; i += blockDim.x)
                    vis[i] = 0u;
            }
        }
        const uint32_t out = sm.ctrl->top_size;
        __syncthreads(); // bitmap is clean and ctrl may be reused
        return out;
    }

    // refine_ (index.hpp:3515-3561) on the ascending list top_d/top_i[0..count): keeps the closest element, then
    // accepts a candidate only if it is not closer to an accepted element than to the centre; because
    // skip_pruned_connections == false (index.hpp:1245) the accepted prefix is followed by whatever sits at
    // positions [accepted, needed) -- stale entries, possibly duplicates.  Destroys the value registers.
    // Returns the size of the resulting view (uniform across the CTA).
    __device__ __forceinline__ uint32_t refine(uint32_t count, uint32_t needed) {
        if (count < needed)
            return count;
        uint32_t submitted = 1, consumed = 1;
        while (submitted < needed && consumed < count) {
            __syncthreads();
            const uint32_t c_id = sm.top_i[consumed] & kIdMask;
            const float c_d = sm.top_d[consumed];
            load_node(c_id);
            for (uint32_t j = threadIdx.x; j < submitted; j += blockDim.x)
                sm.cand_id[j] = sm.top_i[j] & kIdMask;
            __syncthreads();
            eval(submitted);
            __syncthreads();
            bool good = true;
            for (uint32_t j = 0; j < submitted; ++j)
                if (sm.cand_d[j] < c_d) { // index.hpp:3536
                    good = false;
                    break;
                }
            st_dist += submitted;
            if (good) {
                if (threadIdx.x == 0)
                    sm.top_d[submitted] = c_d, sm.top_i[submitted] = sm.top_i[consumed];
                submitted++;
            }
            consumed++;
        }
        __syncthreads();
        return needed; // count >= needed, so shrink(max(submitted, needed)) == needed
    }
};

template <int DM, int SK, int NQ> using Walker = WalkerT<RowEval<DM, SK, NQ>>;
template <int DM> using PqWalker = WalkerT<PqEval<DM>>;

// ---- (DM, SK, NQ) dispatch shared by the launchers ----------------------------------------------------
// PQ indexes: fn(walker type tag) with PqWalker<DM>
template <typename T> struct TypeTag {
    using type = T;
};
template <typename Fn> void dispatch_walker(bool pq, int dm, int sk, int nq, Fn&& fn);

template <typename Fn> void dispatch_walk(int dm, int sk, int nq, Fn&& fn) {
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
    throw CudaError("unsupported metric / scalar kind / dimensionality combination");
}

template <typename Fn> void dispatch_walker(bool pq, int dm, int sk, int nq, Fn&& fn) {
    if (pq) {
        if (dm == DM_L2SQ)
            return fn(TypeTag<PqWalker<DM_L2SQ>>{});
        if (dm == DM_COS)
            return fn(TypeTag<PqWalker<DM_COS>>{});
        throw CudaError("pq index: only l2sq and cos metrics are supported");
    }
    dispatch_walk(dm, sk, nq, [&](auto d, auto s, auto n) {
        fn(TypeTag<Walker<decltype(d)::value, decltype(s)::value, decltype(n)::value>>{});
    });
}

} // namespace lb200
