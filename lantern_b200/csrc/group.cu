// lantern_b200 -- one HNSW graph searched by G GPUs at once ("row-sharded group").
//
// The multi-GPU form of the hot path (index_gt::search -> search_for_one_ -> search_to_find_in_base_,
// U/include/usearch/index.hpp:2680-2730, 3277-3316, 3400-3485).  SURVEY.md 8(e) shards the corpus by contiguous
// row range with one INDEPENDENT graph per GPU and broadcasts every query to every shard; HNSW's cost barely depends on
// N, so that multiplies the total work by G (measured in round 1: 2.7x at 8 GPUs, recall-matched).  Here the corpus is
// still sharded by row range -- GPU r holds the vectors of rows [bounds[r], bounds[r+1]) and nothing else of the 3 KB/row
// payload -- but there is ONE graph: the adjacency lists (4 B per link, 8 % of the corpus at d=768/M=32) are replicated,
// every query is walked once, by one "owner" warp, in the reference's exact order, and each distance is evaluated on the
// GPU that owns the row.  Total row traffic is that of the 1-GPU search, split G ways; the results are those of the 1-GPU
// search on the same graph, id for id (same decision sequence, same fp32 reduction order).
//
// Mapping to the machine.  Every GPU runs the same persistent kernel of independent warps (no CTA-wide barrier anywhere).
// The first O warps are OWNERS: each takes one of this GPU's queries (q mod G == rank) at a time, keeps the query in
// registers, the top list in shared memory, the visited bitmap in HBM, and makes the reference's decisions.  The other H
// warps are a HELPER POOL that measures local rows for remote owners: the (G-1)*O inbound mailboxes are dealt round-robin
// to the helpers, one lane watching one mailbox.  Per expansion the owner sends each rank the ids of the unseen neighbours
// that live there and receives their distances; messages are 8-byte words {payload, flag} written straight into the peer's
// HBM over NVLink (peer-mapped memory; NCCL's "LL" idea: the flag travels with the data, so no fence is needed and one
// NVLink write latency is the whole cost) and polled locally.  Rows are fetched through a per-warp bulk-copy ring
// (cp.async.bulk + mbarrier) behind an L2 prefetch of the rows to come.  The final top-k of a query is stored into every
// GPU's result buffer by its owner (the all-gather of SURVEY 8e, fused into the search epilogue); a last-warp-out flag
// exchange makes kernel completion imply that all peers' results have landed.  No NCCL, no host synchronisation on the hot
// path.  With G = 1 the same kernel is a one-warp-per-query single-GPU search (launch_warp_search).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/lantern_b200.h"
#include "group.h"
#include "walk.cuh"

namespace lb200 {

namespace {

constexpr int kGroupThreads = 128; // 4 independent warps per CTA (no CTA-wide synchronisation anywhere)
constexpr int kGroupWarps = kGroupThreads / 32;

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long pack_word(uint32_t payload, uint32_t flag) {
    return ((unsigned long long)flag << 32) | payload;
}
// 16 bytes of a row that nobody caches on our side (queries live in the root's memory: peer addresses bypass our L2 and
// a stale L1 line of the previous batch must not be served)
__device__ __forceinline__ uint4 ld_nocache_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

struct GroupWarpLayout {
    uint32_t ring, bars, top_d, top_i, cand_id, cand_d, cand_slot, loc, limbo, total;
};
// per warp: R row slots (the bulk-copy ring), R mbarriers, the top list, the candidate arrays
__host__ __device__ inline GroupWarpLayout group_warp_layout(uint32_t row_bytes, uint32_t R, uint32_t L, uint32_t cap) {
    GroupWarpLayout l;
    uint32_t o = 0;
    l.ring = o, o += R * ((row_bytes + 15u) & ~15u);
    l.bars = o, o += 8 * R;
    l.top_d = o, o += 4 * L;
    l.top_i = o, o += 4 * L;
    l.cand_id = o, o += 4 * cap;
    l.cand_d = o, o += 4 * cap;
    l.cand_slot = o, o += (2 * cap + 3u) & ~3u;
    l.loc = o, o += (cap + 3u) & ~3u;
    l.limbo = o, o += 4 * kLimboCap;
    l.total = (o + 15u) & ~15u;
    return l;
}
// rows in flight per warp: 2 for 3 KB rows, more for narrow rows (about 6 KB of staging per warp)
inline uint32_t group_ring_slots(uint32_t row_bytes) {
    uint32_t r = 6144u / row_bytes;
    r = r < 2 ? 2 : (r > 8 ? 8 : r);
    if (const char* e = getenv("LB200_GROUP_RING_SLOTS")) // experiments: deeper rings cost resident warps
        if (atoi(e) >= 1 && atoi(e) <= 8)
            r = (uint32_t)atoi(e);
    return r;
}

// ---- one warp = one query slot -------------------------------------------------------------------------------------
// Header payload of the owner -> helper mailbox: count (bits 0-8, 1..256 ids follow) | query index << 9, or kMsgExit (the
// launch is over for this slot).  The query index travels with EVERY request -- the header is a single word that the next
// message overwrites, so a separate "start of query" message could be lost before the helper had seen it; the helper loads a
// query into shared memory when the index in a request differs from the one it holds.
constexpr uint32_t kMsgExit = 0xFFFFFFFFu;
constexpr uint32_t kMsgCountBits = 9;
constexpr size_t kGroupMaxBatch = (size_t)1 << (32 - kMsgCountBits - 1);

template <int DM, int SK, int NQ> struct GroupWarp {
    const GroupLaunch& p;
    int lane;
    uint32_t slot, nchunks;
    uint8_t* ws; // this warp's shared memory
    uint4 qreg[NQ]; // the query, in registers (lane l holds chunks l, l+32, ...)
    uint32_t phase_bits; // parity of each ring slot's mbarrier
    float a2;
    uint32_t seq;   // owner: messages sent by this slot in this launch
    uint32_t last;  // helper: flag of the last header seen
    uint32_t cur_q; // the query this slot holds in shared memory (0xFFFFFFFF: none)
    bool dead, waited;
    unsigned long long t0;
    uint32_t st_dist, st_pops, st_hops, st_rounds, st_rows;
    // owner: SM cycles spent producing a round's ids, measuring the local rows, waiting for the helpers, consuming
    uint32_t cy_produce, cy_local, cy_wait, cy_consume; // (32 bits: 2 s of cycles per warp and launch)

    __device__ __forceinline__ explicit GroupWarp(const GroupLaunch& gp) : p(gp) {}

    // shared-memory arrays are addressed from one base (fewer live registers than eight pointers)
    __device__ __forceinline__ GroupWarpLayout lay() const { return group_warp_layout(p.g.row_bytes, p.ring_slots, p.L, p.cap); }
    __device__ __forceinline__ uint8_t* ring() const { return ws; }
    __device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(ws + lay().bars); }
    __device__ __forceinline__ float* top_d() const { return reinterpret_cast<float*>(ws + lay().top_d); }
    __device__ __forceinline__ uint32_t* top_i() const { return reinterpret_cast<uint32_t*>(ws + lay().top_i); }
    __device__ __forceinline__ uint32_t* cand_id() const { return reinterpret_cast<uint32_t*>(ws + lay().cand_id); }
    __device__ __forceinline__ float* cand_d() const { return reinterpret_cast<float*>(ws + lay().cand_d); }
    __device__ __forceinline__ uint16_t* cand_slot() const { return reinterpret_cast<uint16_t*>(ws + lay().cand_slot); }
    __device__ __forceinline__ uint8_t* loc() const { return ws + lay().loc; }
    __device__ __forceinline__ uint32_t* limbo() const { return reinterpret_cast<uint32_t*>(ws + lay().limbo); }

    __device__ __noinline__ bool timed_out() {
        if (ld_sys_u32(p.err[p.me]) != 0u)
            return true;
        if (globaltimer_ns() - t0 > p.timeout_ns) {
            for (uint32_t r = 0; r < p.G; ++r)
                st_sys_u32(p.err[r], 1u + p.me);
            return true;
        }
        return false;
    }
    // spin until the word carries `flag`; returns its payload
    __device__ __forceinline__ uint32_t wait_word(const unsigned long long* addr, uint32_t flag) {
        uint32_t spins = 0;
        for (;;) {
            const unsigned long long v = ld_sys_u64(addr);
            if ((uint32_t)(v >> 32) == flag)
                return (uint32_t)v;
            if (dead)
                return 0u;
            if ((++spins & 2047u) == 0u && timed_out()) {
                dead = true;
                return 0u;
            }
        }
    }

    // non-root ranks read the queries from the root's staging buffer once the root's kernel has announced them
    __device__ __forceinline__ void wait_queries() {
        if (waited)
            return;
        if (lane == 0) {
            uint32_t spins = 0;
            while (ld_sys_u64(p.qready[p.me]) != (unsigned long long)p.epoch) {
                if ((++spins & 2047u) == 0u && timed_out()) {
                    dead = true;
                    break;
                }
                __nanosleep(64);
            }
        }
        dead = __any_sync(0xffffffffu, dead);
        waited = true;
    }

    __device__ __forceinline__ void load_query(uint32_t q) {
        wait_queries();
        const uint4* src = reinterpret_cast<const uint4*>(p.queries + (size_t)q * p.query_stride);
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const uint32_t c = lane + 32 * i;
            qreg[i] = c < nchunks ? ld_nocache_u4(src + c) : make_uint4(0, 0, 0, 0);
            part = norm_add(part, query_norm_chunk<DM, SK>(qreg[i]));
        }
        a2 = 0.f;
        if constexpr (DM == DM_COS)
            a2 = warp_sum(part);
    }

    // ---- distances query -> a list of LOCAL rows, through this warp's bulk-copy ring (cp.async.bulk + mbarrier, as
    // RowEval in walk.cuh but private to the warp): ring_slots rows are in flight at once, a slot is refilled as soon as it
    // has been read.  OWNER: entry t of the list is candidate loc[t]; helper: entry t is cand_id[t].  Lane layout and
    // reduction order are RowEval::row_distance's, so every value is bit-identical to the 1-GPU kernel's.
    __device__ __forceinline__ void issue_row(uint32_t s, uint32_t id) {
        uint64_t* bar = bars() + s;
        mbar_arrive_expect_tx(bar, p.g.row_bytes);
        bulk_g2s(ring() + (size_t)s * p.g.row_bytes, p.g.vectors + (size_t)(id - p.bounds[p.me]) * p.g.row_bytes, p.g.row_bytes, bar);
    }
    template <bool OWNER> __device__ __forceinline__ void eval_local(uint32_t n) {
        const uint32_t R = p.ring_slots;
        const uint32_t* ids = cand_id();
        const uint8_t* lj = loc();
        // The ring keeps R rows in flight; the rows behind them are already on their way to L2 (one prefetch per 128-byte line,
        // kPrefetchRows rows ahead), so a slot's refill is an L2 hit instead of a DRAM round trip.  With G GPUs a warp's share of
        // an expansion is a handful of rows: all of them are requested from DRAM at once.
        // Only for SHORT lists (the latency-bound regime of 4+ GPUs): with 24-48 rows per list (1-2 GPUs) the kernel is
        // HBM-bound and 8 rows x 3 KB x 3500 warps of prefetch would not fit in L2 -- measured: 20.9 instead of 15.9 ms at G = 1.
        constexpr uint32_t kPrefetchRows = 8, kPrefetchMaxList = 16;
        const uint32_t pf_off = (uint32_t)lane * 128u;
        const bool pf_lane = pf_off < p.g.row_bytes && n <= kPrefetchMaxList;
        const uint8_t* base = p.g.vectors - (size_t)p.bounds[p.me] * p.g.row_bytes;
        for (uint32_t t = R; t < min(n, R + kPrefetchRows); ++t)
            if (pf_lane)
                prefetch_l2(base + (size_t)ids[OWNER ? lj[t] : t] * p.g.row_bytes + pf_off);
        if ((uint32_t)lane < min(n, R))
            issue_row(lane, ids[OWNER ? lj[lane] : lane]);
        uint32_t s = 0;
#pragma unroll 1
        for (uint32_t t = 0; t < n; ++t) {
            if (t + R + kPrefetchRows < n && pf_lane)
                prefetch_l2(base + (size_t)ids[OWNER ? lj[t + R + kPrefetchRows] : t + R + kPrefetchRows] * p.g.row_bytes + pf_off);
            mbar_wait(bars() + s, (phase_bits >> s) & 1u);
            phase_bits ^= 1u << s;
            const uint4* row = reinterpret_cast<const uint4*>(ring() + (size_t)s * p.g.row_bytes);
            DistAcc<DM, SK> acc;
            acc.reset();
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const uint32_t c = lane + 32 * i;
                if (c < nchunks)
                    accum_chunk<DM, SK>(acc, qreg[i], row[c]);
            }
            const float d = finish_distance<DM, SK>(acc, a2);
            if (lane == 0)
                cand_d()[OWNER ? lj[t] : t] = d;
            __syncwarp();
            if (t + R < n && lane == 0) {
                fence_proxy_async(); // our generic-proxy reads of the slot precede the async refill
                issue_row(s, ids[OWNER ? lj[t + R] : t + R]);
            }
            s = (s + 1 == R) ? 0 : s + 1;
        }
        st_rows += n;
    }

    // ---- owner: distances query -> cand_id[0..n) into cand_d[0..n), each evaluated where the row lives -----------
    __device__ __forceinline__ void eval_round(uint32_t n) {
        const uint32_t G = p.G, me = p.me, cap = p.cap;
        const uint32_t flag = p.flag_base + (++seq);
        const uint32_t lt = (1u << lane) - 1u;
        uint32_t cnt = 0; // lane d < G: ids addressed to rank d in this round
#pragma unroll 1
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t j = base + lane;
            const bool valid = j < n;
            const uint32_t id = valid ? cand_id()[j] : 0u;
            uint32_t dst = 0xFFu;
            if (valid) {
                dst = 0;
                while (id >= p.bounds[dst + 1])
                    ++dst;
            }
            uint32_t pos = 0;
#pragma unroll 1
            for (uint32_t d = 0; d < G; ++d) {
                const uint32_t m = __ballot_sync(0xffffffffu, dst == d);
                const uint32_t c_d = __shfl_sync(0xffffffffu, cnt, d);
                if (dst == d)
                    pos = c_d + __popc(m & lt);
                if ((uint32_t)lane == d)
                    cnt += __popc(m);
            }
            if (valid) {
                cand_slot()[j] = (uint16_t)((dst << 8) | pos);
                if (dst != me)
                    st_sys_u64(p.req[dst] + ((size_t)me * p.O + slot) * (1 + cap) + 1 + pos, pack_word(id, flag));
                else
                    loc()[pos] = (uint8_t)j;
            }
        }
        if ((uint32_t)lane < G && (uint32_t)lane != me && cnt)
            st_sys_u64(p.req[lane] + ((size_t)me * p.O + slot) * (1 + cap), pack_word(cnt | (cur_q << kMsgCountBits), flag));
        const uint32_t nloc = __shfl_sync(0xffffffffu, cnt, me);
        __syncwarp();
        const uint32_t c0 = (uint32_t)clock64();
        eval_local<true>(nloc);
        const uint32_t c1 = (uint32_t)clock64();
        cy_local += c1 - c0;
        const unsigned long long* inbox = p.resp[me] + (size_t)slot * G * cap;
#pragma unroll 1
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t j = base + lane;
            if (j < n) {
                const uint32_t s = cand_slot()[j], dst = s >> 8, pos = s & 255u;
                if (dst != me)
                    cand_d()[j] = __uint_as_float(wait_word(inbox + (size_t)dst * cap + pos, flag));
            }
        }
        dead = __any_sync(0xffffffffu, dead);
        cy_wait += (uint32_t)clock64() - c1;
        st_rounds += 1;
        __syncwarp();
    }

    // ---- helper: one of the H warps of this GPU that measure local rows for REMOTE owners.  The (G-1) * O inbound
    // mailboxes (one per remote owner slot) are dealt round-robin to the helpers; lane i of helper h watches mailbox
    // h + i * H, so one volatile load per lane polls the helper's whole set.  A request names its query; the query vector
    // sits next to the mailbox (the owner stored it there, fence, before its first request) and is loaded into registers
    // whenever the helper switches mailbox or the mailbox switches query. --------------------------------------------------
    __device__ __forceinline__ void load_mailbox_query(const uint8_t* src_bytes) {
        const uint4* src = reinterpret_cast<const uint4*>(src_bytes);
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const uint32_t c = lane + 32 * i;
            qreg[i] = c < nchunks ? ld_nocache_u4(src + c) : make_uint4(0, 0, 0, 0);
            part = norm_add(part, query_norm_chunk<DM, SK>(qreg[i]));
        }
        a2 = 0.f;
        if constexpr (DM == DM_COS)
            a2 = warp_sum(part);
    }

    __device__ __forceinline__ void run_helper(uint32_t h) {
        const uint32_t cap = p.cap, G = p.G, O = p.O, H = p.H, M = (G - 1) * O;
        const uint32_t mp = h + (uint32_t)lane * H; // this lane's mailbox among the M remote ones (if < M)
        bool alive = mp < M;
        uint32_t my_last = p.flag_base;
        uint32_t src = 0, oslot = 0;
        if (alive) {
            src = mp / O, oslot = mp - src * O;
            if (src >= p.me)
                ++src; // own mailboxes are skipped
        }
        const unsigned long long* my_hdr = p.req[p.me] + ((size_t)src * O + oslot) * (1 + cap);
        uint32_t cur_m = 0xFFFFFFFFu;
        uint32_t spins = 0;
#pragma unroll 1
        while (__any_sync(0xffffffffu, alive)) {
            uint32_t pay = 0, flag = 0;
            bool fresh = false;
            if (alive) {
                const unsigned long long v = ld_sys_u64(my_hdr);
                flag = (uint32_t)(v >> 32), pay = (uint32_t)v;
                const uint32_t ahead = flag - my_last;
                fresh = ahead != 0u && ahead < (1u << 20) && flag - p.flag_base < (1u << 20);
            }
            uint32_t ready = __ballot_sync(0xffffffffu, fresh);
            if (!ready) {
                if ((++spins & 1023u) == 0u) {
                    if (lane == 0 && timed_out())
                        dead = true;
                    dead = __any_sync(0xffffffffu, dead);
                    if (dead)
                        return;
                }
                __nanosleep(32);
                continue;
            }
            spins = 0;
            while (ready) {
                const int b = __ffs(ready) - 1;
                ready &= ready - 1;
                const uint32_t r_pay = __shfl_sync(0xffffffffu, pay, b), r_flag = __shfl_sync(0xffffffffu, flag, b);
                const uint32_t r_src = __shfl_sync(0xffffffffu, src, b), r_oslot = __shfl_sync(0xffffffffu, oslot, b);
                if (lane == b)
                    my_last = r_flag;
                if (r_pay == kMsgExit) {
                    if (lane == b)
                        alive = false;
                    continue;
                }
                const uint32_t q = r_pay >> kMsgCountBits, n_d = r_pay & ((1u << kMsgCountBits) - 1u);
                const size_t m = (size_t)r_src * O + r_oslot;
                if ((uint32_t)m != cur_m || q != cur_q) {
                    load_mailbox_query(p.reqq[p.me] + m * p.g.row_bytes);
                    cur_m = (uint32_t)m, cur_q = q;
                }
                const unsigned long long* hdr = p.req[p.me] + m * (1 + cap);
                unsigned long long* outbox = p.resp[r_src] + ((size_t)r_oslot * G + p.me) * cap;
#pragma unroll 1
                for (uint32_t base = 0; base < n_d; base += 32) {
                    const uint32_t cnt = min(32u, n_d - base);
                    const uint32_t my_id = (uint32_t)lane < cnt ? wait_word(hdr + 1 + base + lane, r_flag) : 0u;
                    dead = __any_sync(0xffffffffu, dead);
                    if (dead)
                        return;
                    if ((uint32_t)lane < cnt)
                        cand_id()[lane] = my_id;
                    __syncwarp();
                    eval_local<false>(cnt);
                    __syncwarp();
                    if ((uint32_t)lane < cnt)
                        st_sys_u64(outbox + base + lane, pack_word(__float_as_uint(cand_d()[lane]), r_flag));
                    __syncwarp();
                }
            }
        }
    }

    __device__ __forceinline__ void tell_helpers(uint32_t payload) {
        const uint32_t flag = p.flag_base + (++seq);
        if ((uint32_t)lane < p.G && (uint32_t)lane != p.me)
            st_sys_u64(p.req[lane] + ((size_t)p.me * p.O + slot) * (1 + p.cap), pack_word(payload, flag));
    }

    // ---- owner: the reference's walk of ONE query by one warp.  A single produce -> evaluate -> consume loop serves the entry
    // point, the greedy descent (search_for_one_, index.hpp:3277-3316) and the base-layer beam (search_to_find_in_base_,
    // :3400-3485), so that eval_round is instantiated once. -------------------------------------------------------------
    __device__ __forceinline__ void run_owner(uint32_t q) {
        const uint32_t M0 = p.g.M0, L = p.L;
        uint32_t* vis = p.vis + (size_t)slot * p.words_per_slot;
        uint32_t* touched = p.touched + (size_t)slot * p.touched_cap;
        load_query(q);
        cur_q = q;
        if (p.G > 1) { // the helpers read the query from the mailbox of this owner slot on THEIR GPU
            for (uint32_t d = 0; d < p.G; ++d) {
                if (d == p.me)
                    continue;
                uint4* dst = reinterpret_cast<uint4*>(p.reqq[d] + ((size_t)p.me * p.O + slot) * p.g.row_bytes);
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    const uint32_t c = lane + 32 * i;
                    if (c < nchunks)
                        dst[c] = qreg[i];
                }
            }
            __threadfence_system(); // the vector is in place before the first request that names this query can be seen
        }
        int level = -1; // -1: measuring the entry point; >= 1: greedy on that level; 0: beam on the base layer
        uint32_t cur = p.g.entry;
        float cur_d = 0.f;
        uint32_t size = 0, cursor = 0, ntouched = 0, limbo_n = 0;
        float limbo_d = 0.f;
#pragma unroll 1
        while (!dead) {
            // ---- produce the ids to measure ----
            const uint32_t cp0 = (uint32_t)clock64();
            uint32_t n = 0;
            if (level < 0) {
                if (lane == 0)
                    cand_id()[0] = cur;
                n = 1;
            } else if (level > 0) {
                const uint32_t* list = p.g.upper_adj + ((size_t)__ldg(p.g.upper_ref + cur) + (level - 1)) * p.g.M;
                for (uint32_t off = 0; off < p.g.M; off += 32) {
                    const uint32_t id = (off + lane < p.g.M) ? __ldg(list + off + lane) : kNoNeighbor;
                    const bool valid = id != kNoNeighbor;
                    const uint32_t m = __ballot_sync(0xffffffffu, valid);
                    if (valid)
                        cand_id()[n + __popc(m & ((1u << lane) - 1u))] = id;
                    n += __popc(m);
                }
                st_hops += 1;
            } else {
                // pop the closest unexpanded entry (walk.cuh TopSmem::pop), or a tie waiting in limbo
                uint32_t c = kNoNeighbor;
                uint32_t* ti = top_i();
                if (cursor < size) {
                    c = ti[cursor];
                    __syncwarp();
                    if (lane == 0)
                        ti[cursor] = c | kExpandedBit;
                    __syncwarp();
                    uint32_t nxt = size;
                    for (uint32_t b = cursor + 1; b < size; b += 32) {
                        const uint32_t e = b + lane;
                        const bool un = e < size && !(ti[e] & kExpandedBit);
                        const uint32_t m = __ballot_sync(0xffffffffu, un);
                        if (m) {
                            nxt = b + __ffs(m) - 1;
                            break;
                        }
                    }
                    cursor = nxt;
                } else if (limbo_n) {
                    c = limbo()[--limbo_n];
                }
                if (c == kNoNeighbor)
                    break; // the beam is exhausted: index.hpp:3445 / queue empty
                const uint32_t* list = p.g.adj0 + (size_t)c * M0;
                if (M0 == 64) {
                    // both halves of a 64-wide list at once: the two bitmap atomics of a lane are in flight together (one L2
                    // atomic round trip per expansion instead of two).  Duplicates are legal in reference graphs (refine_
                    // padding, index.hpp:3554-3558): only the FIRST occurrence in stored order may count as unseen, so the
                    // second half is also checked against the first (through the candidate array, free at this point).
                    const uint32_t id0 = __ldg(list + lane), id1 = __ldg(list + 32 + lane);
                    const bool v0 = id0 != kNoNeighbor, v1 = id1 != kNoNeighbor;
                    cand_id()[lane] = id0;
                    __syncwarp();
                    const uint32_t p0 = __match_any_sync(0xffffffffu, id0), p1 = __match_any_sync(0xffffffffu, id1);
                    bool first0 = v0 && ((uint32_t)(__ffs(p0) - 1) == (uint32_t)lane);
                    bool first1 = v1 && ((uint32_t)(__ffs(p1) - 1) == (uint32_t)lane);
                    if (first1)
                        for (int t = 0; t < 32; ++t)
                            first1 &= cand_id()[t] != id1;
                    __syncwarp();
                    bool fresh0 = false, fresh1 = false;
                    uint32_t old0 = 0, old1 = 0;
                    const uint32_t bit0 = 1u << (id0 & 31), bit1 = 1u << (id1 & 31);
                    if (first0)
                        old0 = atomicOr(&vis[id0 >> 5], bit0);
                    if (first1)
                        old1 = atomicOr(&vis[id1 >> 5], bit1);
                    fresh0 = first0 && !(old0 & bit0), fresh1 = first1 && !(old1 & bit1);
                    const uint32_t m0 = __ballot_sync(0xffffffffu, fresh0), m1 = __ballot_sync(0xffffffffu, fresh1);
                    const uint32_t lt = (1u << lane) - 1u, n0 = __popc(m0);
                    if (fresh0) {
                        const uint32_t r = __popc(m0 & lt);
                        cand_id()[r] = id0;
                        if (ntouched + r < p.touched_cap)
                            touched[ntouched + r] = id0 >> 5;
                    }
                    __syncwarp(); // (the compaction above overwrites entries another lane compared against: all compares are done)
                    if (fresh1) {
                        const uint32_t r = n0 + __popc(m1 & lt);
                        cand_id()[r] = id1;
                        if (ntouched + r < p.touched_cap)
                            touched[ntouched + r] = id1 >> 5;
                    }
                    n = n0 + __popc(m1);
                    ntouched += n;
                } else
                    for (uint32_t off = 0; off < M0; off += 32) {
                        const uint32_t id = (off + lane < M0) ? __ldg(list + off + lane) : kNoNeighbor;
                        const bool valid = id != kNoNeighbor;
                        if (!__any_sync(0xffffffffu, valid))
                            break;
                        // duplicate ids inside one list are legal in reference graphs (refine_ padding, index.hpp:3554-3558)
                        const uint32_t peers = __match_any_sync(0xffffffffu, id);
                        const bool first = valid && ((uint32_t)(__ffs(peers) - 1) == (uint32_t)lane);
                        bool fresh = false;
                        if (first) {
                            const uint32_t bit = 1u << (id & 31);
                            fresh = !(atomicOr(&vis[id >> 5], bit) & bit);
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, fresh);
                        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                        if (fresh) {
                            cand_id()[n + rank] = id;
                            if (ntouched + rank < p.touched_cap)
                                touched[ntouched + rank] = id >> 5;
                        }
                        n += __popc(m);
                        ntouched += __popc(m);
                    }
                st_pops += 1;
            }
            __syncwarp();
            const uint32_t cp1 = (uint32_t)clock64();
            // ---- evaluate: every id on the GPU that holds its row ----
            if (n)
                eval_round(n);
            if (dead)
                break;
            const uint32_t cp2 = (uint32_t)clock64();
            cy_produce += cp1 - cp0; // classification + sends are counted with the evaluation
            st_dist += n;
            // ---- consume ----
            if (level < 0) {
                cur_d = cand_d()[0];
                level = p.g.max_level;
            } else if (level > 0) {
                float best = cur_d; // one pass of index.hpp:3304-3311 == first minimum below cur_d
                int bi = -1;
                for (uint32_t j = 0; j < n; ++j) {
                    const float d = cand_d()[j];
                    if (d < best)
                        best = d, bi = (int)j;
                }
                if (bi >= 0)
                    cur = cand_id()[bi], cur_d = best;
                else
                    --level;
                __syncwarp();
            } else {
                // index.hpp:3470: accepted iff top.size() < L || d < radius; parallel pre-filter, then replay in stored order
                float* td = top_d();
                uint32_t* ti = top_i();
                for (uint32_t base = 0; base < n; base += 32) {
                    const uint32_t j = base + lane;
                    const float dj = j < n ? cand_d()[j] : INFINITY;
                    float radius = td[size - 1];
                    uint32_t m = __ballot_sync(0xffffffffu, j < n && (size < L || dj < radius));
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const float d = __shfl_sync(0xffffffffu, dj, b);
                        if (size < L || d < radius) {
                            const uint32_t id = cand_id()[base + b];
                            float ev_d;
                            uint32_t ev_i;
                            top_insert(td, ti, size, cursor, L, d, id, lane, ev_d, ev_i);
                            if ((p.g.flags & 2u) && lane == 0)
                                prefetch_l2(p.g.adj0 + (size_t)id * M0);
                            radius = td[size - 1];
                            if (limbo_n && radius < limbo_d)
                                limbo_n = 0;
                            if (ev_i != kNoNeighbor && !(ev_i & kExpandedBit) && ev_d == radius) {
                                if (limbo_n < kLimboCap) {
                                    if (lane == 0)
                                        limbo()[limbo_n] = ev_i;
                                    limbo_n++, limbo_d = radius;
                                    __syncwarp();
                                } else if (lane == 0) {
                                    atomicAdd(&p.counters[7], 1ull); // rare by construction
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            cy_consume += (uint32_t)clock64() - cp2;
            if (level == 0 && size == 0) { // the descent is over: open the beam at `cur` (its distance is known; the
                if (lane == 0) {           // reference measures it again, index.hpp:3436, so the counter advances)
                    top_d()[0] = cur_d, top_i()[0] = cur;
                    atomicOr(&vis[cur >> 5], 1u << (cur & 31));
                    touched[0] = cur >> 5;
                }
                size = 1, cursor = 0, ntouched = 1;
                st_dist += 1;
                __syncwarp();
            }
        }
        // un-visit only the words this walk touched
        if (ntouched <= p.touched_cap) {
            for (uint32_t i = lane; i < ntouched; i += 32)
                vis[touched[i]] = 0u;
        } else {
            for (size_t i = lane; i < p.words_per_slot; i += 32)
                vis[i] = 0u;
        }
        // results -> every rank (the all-gather of SURVEY 8e, fused); top is ascending, shrink(k), keys
        const uint32_t found = dead ? 0u : min(size, p.k);
        for (uint32_t i = lane; i < p.k; i += 32) {
            unsigned long long key = ~0ull;
            float d = INFINITY;
            if (i < found) {
                key = __ldg(p.g.keys + (top_i()[i] & kIdMask));
                d = top_d()[i];
            }
            for (uint32_t r = 0; r < p.G; ++r) {
                p.res_keys[r][(size_t)q * p.k + i] = key;
                p.res_dists[r][(size_t)q * p.k + i] = d;
            }
        }
        if ((uint32_t)lane < p.G)
            p.res_counts[lane][q] = found;
        __syncwarp();
    }
};

template <int DM, int SK, int NQ>
__global__ void __launch_bounds__(kGroupThreads, 6) group_search_kernel(const __grid_constant__ GroupLaunch p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    GroupWarp<DM, SK, NQ> w(p);
    w.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    w.slot = blockIdx.x * kGroupWarps + warp;
    w.nchunks = p.g.row_bytes / 16;
    w.ws = smem_raw + (size_t)warp * group_warp_layout(p.g.row_bytes, p.ring_slots, p.L, p.cap).total;
    w.phase_bits = 0;
    if (w.lane == 0) {
        for (uint32_t i = 0; i < p.ring_slots; ++i)
            mbar_init(w.bars() + i, 1);
        fence_mbar_init();
    }
    __syncwarp();
    w.seq = 0, w.last = p.flag_base, w.dead = false, w.a2 = 0.f, w.cur_q = 0xFFFFFFFFu;
    w.waited = (p.me == p.root);
    w.st_dist = w.st_pops = w.st_hops = w.st_rounds = w.st_rows = 0;
    w.cy_produce = w.cy_local = w.cy_wait = w.cy_consume = 0;
    w.t0 = globaltimer_ns();

    // the root's query staging buffer is complete when its kernel starts (stream order): tell everybody
    if (p.me == p.root && w.slot == 0 && (uint32_t)w.lane < p.G)
        st_sys_u64(p.qready[w.lane], (unsigned long long)p.epoch);
    if (w.slot < p.O) {
        // owner warp: this GPU's queries (q mod G == me) are handed to its owner warps as they become free
        for (;;) {
            uint32_t idx = 0;
            if (w.lane == 0)
                idx = (uint32_t)atomicAdd(&p.counters[6], 1ull);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            const unsigned long long q = (unsigned long long)idx * p.G + p.me;
            if (q >= p.nq || w.dead)
                break;
            w.run_owner((uint32_t)q);
        }
        if (p.G > 1)
            w.tell_helpers(kMsgExit);
    } else {
        w.run_helper(w.slot - p.O);
    }
    // completion: results of this GPU's owners are visible system-wide before its done flag is
    __threadfence_system();
    __syncwarp();
    bool last_warp = false;
    if (w.lane == 0) {
        atomicAdd(&p.counters[1], (unsigned long long)w.st_dist);
        atomicAdd(&p.counters[2], (unsigned long long)w.st_pops);
        atomicAdd(&p.counters[3], (unsigned long long)w.st_hops);
        atomicAdd(&p.counters[4], (unsigned long long)w.st_rounds);
        atomicAdd(&p.counters[5], (unsigned long long)w.st_rows);
        if (w.cy_produce | w.cy_wait) {
            atomicAdd(&p.counters[8], (unsigned long long)w.cy_produce), atomicAdd(&p.counters[9], (unsigned long long)w.cy_local);
            atomicAdd(&p.counters[10], (unsigned long long)w.cy_wait), atomicAdd(&p.counters[11], (unsigned long long)w.cy_consume);
        }
        __threadfence();
        last_warp = atomicAdd(&p.counters[0], 1ull) == (unsigned long long)gridDim.x * kGroupWarps - 1ull;
    }
    last_warp = __shfl_sync(0xffffffffu, last_warp, 0);
    if (last_warp) {
        __threadfence_system();
        if ((uint32_t)w.lane < p.G) {
            st_sys_u64(p.done[w.lane] + p.me, (unsigned long long)p.epoch);
            uint32_t spins = 0;
            while (ld_sys_u64(p.done[p.me] + w.lane) != (unsigned long long)p.epoch) {
                if ((++spins & 2047u) == 0u && w.timed_out())
                    break;
                __nanosleep(64);
            }
        }
        __threadfence_system();
    }
}

template <int DM, int SK, int NQ> void group_launch_one(const GroupLaunch& p, uint32_t grid, size_t smem, cudaStream_t stream) {
    auto kern = group_search_kernel<DM, SK, NQ>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kGroupThreads, smem, stream>>>(p);
    LB_CUDA(cudaGetLastError());
    count_launch();
}
template <int DM, int SK, int NQ> int group_occupancy_one(size_t smem) {
    auto kern = group_search_kernel<DM, SK, NQ>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = 0;
    LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, kGroupThreads, smem));
    return blocks;
}

// in-degree of the base layer per bucket of consecutive rows: how often the rows of a bucket appear in adjacency lists is
// how often they will be measured.  Row ranges of equal WEIGHT (not equal length) balance the ranks: early rows of an HNSW
// build collect more links than late ones, so equal-length ranges would leave the first rank with ~40 % more work.
constexpr uint32_t kBalanceBuckets = 4096;
__global__ void indegree_hist_kernel(const uint32_t* __restrict__ adj0, size_t total, uint32_t n, unsigned long long* __restrict__ hist) {
    __shared__ uint32_t sh[kBalanceBuckets];
    for (uint32_t i = threadIdx.x; i < kBalanceBuckets; i += blockDim.x)
        sh[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t id = adj0[i];
        if (id < n)
            atomicAdd(&sh[(uint32_t)(((unsigned long long)id * kBalanceBuckets) / n)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kBalanceBuckets; i += blockDim.x)
        if (sh[i])
            atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

__global__ void group_copy_results_kernel(const uint64_t* __restrict__ rk, const float* __restrict__ rd, const uint32_t* __restrict__ rc,
                                          uint64_t* __restrict__ keys, float* __restrict__ dists, uint32_t* __restrict__ counts,
                                          size_t nq, size_t k) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * k) {
        if (keys)
            keys[i] = rk[i];
        if (dists)
            dists[i] = rd[i];
    }
    if (counts && i < nq)
        counts[i] = rc[i];
}

} // namespace

// =====================================================================================================================
// host side
// =====================================================================================================================

// Roles of a launch: O owner warps (one query each at a time; enough for every query this rank owns, but leaving a quarter of
// the resident warps to the helpers) and H helper warps that share the (G-1) * O inbound mailboxes -- no more helpers than
// mailboxes, and no helper watches more than 32 (one per lane).  Deterministic in (world, nq, W, Omax): every rank computes
// the same plan.  Returns false when the resident warps cannot cover the mailboxes.
bool group_plan(int world, size_t nq, uint32_t W, uint32_t Omax, uint32_t& O, uint32_t& H) {
    O = (uint32_t)round_up((nq + world - 1) / world, kGroupWarps), H = 0;
    O = std::min(O, Omax);
    if (world > 1) {
        O = std::min<uint32_t>(O, (W - W / 4) & ~3u);
        if (O == 0)
            return false;
        const uint32_t M = (uint32_t)(world - 1) * O;
        H = std::min<uint32_t>(W - O, (uint32_t)round_up(M, kGroupWarps));
        H = std::max<uint32_t>(H, (uint32_t)round_up((M + 31) / 32, kGroupWarps));
        return O + H <= W;
    }
    O = std::min(O, W & ~3u);
    return O > 0;
}

// every spin in the kernel gives up after this long (LB200_GROUP_TIMEOUT_S, default 20 s; raise it under compute-sanitizer)
static unsigned long long group_timeout_ns() {
    double sec = 20.0;
    if (const char* e = getenv("LB200_GROUP_TIMEOUT_S"))
        sec = atof(e) > 0 ? atof(e) : sec;
    return (unsigned long long)(sec * 1e9);
}

struct DeviceGuard { // the group switches devices; the caller's current device is put back when an entry point returns
    int dev = 0;
    DeviceGuard() { cudaGetDevice(&dev); }
    ~DeviceGuard() { cudaSetDevice(dev); }
};

struct PtrBlob { // how one device allocation is handed to the other ranks
    cudaIpcMemHandle_t ipc;
    void* raw;
};

struct GroupConfigBlob {
    uint32_t valid;
    int32_t metric_kind, scalar_kind;
    uint64_t dims, M, M0, ef, n, upper_lists, row_bytes, vec_bytes;
    uint32_t entry;
    int32_t max_level;
    uint32_t flags;
    uint32_t bounds[kGroupMax + 1]; // row ranges, balanced by base-layer in-degree
    PtrBlob vectors, adj0, upper_ref, upper_adj, keys;
};

struct SlabLayout {
    size_t req, reqq, resp, res_keys, res_dists, res_counts, done, qready, err, qbuf, total;
};

// Omax = owner slots per rank the mailboxes are sized for
static SlabLayout slab_layout(uint32_t Omax, uint32_t G, uint32_t cap, size_t res_cap, size_t max_batch, size_t qrow) {
    SlabLayout s;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += round_up(bytes, 256);
        return at;
    };
    s.req = take((size_t)G * Omax * (1 + cap) * 8);
    s.resp = take((size_t)Omax * G * cap * 8);
    s.reqq = take((size_t)G * Omax * qrow); // (after resp: renew_flags clears [req, reqq) only)
    s.res_keys = take(res_cap * 8);
    s.res_dists = take(res_cap * 4);
    s.res_counts = take(max_batch * 4);
    s.done = take(kGroupMax * 8);
    s.qready = take(8);
    s.err = take(4);
    s.qbuf = take(max_batch * qrow);
    s.total = o;
    return s;
}

class GroupRank {
  public:
    int rank = 0, world = 1, device = 0;
    bool local = false; // all ranks live in this process (peer pointers are used directly)
    IndexConfig cfg;
    int dist_mode = 0;
    size_t n = 0, row_bytes = 0, vec_bytes = 0, upper_lists = 0;
    uint32_t entry = 0, flags = 2;
    int32_t max_level = -1;
    uint32_t bounds[kGroupMax + 1] = {0};
    // local copies
    uint8_t* d_rows = nullptr;
    uint32_t *d_adj0 = nullptr, *d_upper_ref = nullptr, *d_upper_adj = nullptr;
    uint64_t* d_keys = nullptr;
    // mailboxes
    uint8_t* slab = nullptr;
    uint8_t* peer_slab[kGroupMax] = {nullptr};
    SlabLayout lay{};
    uint32_t Wmax = 0, Omax = 0, cap = 0;
    size_t res_cap = 0, max_batch = 0;
    // owner scratch
    uint32_t *d_vis = nullptr, *d_touched = nullptr;
    size_t words_per_slot = 0;
    uint32_t touched_cap = 16384;
    unsigned long long* d_counters = nullptr;
    uint32_t epoch = 0;
    std::map<uint32_t, uint32_t> W_for_L; // agreed resident warps per GPU for a beam width
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t last_nq = 0;
    std::vector<void*> opened; // IPC mappings to close
    void* d_in = nullptr;      // root: staging of the raw input queries of the host-buffer entry point
    size_t in_bytes = 0;
    std::vector<uint32_t> h_counts;
    uint32_t h_err = 0;

    ~GroupRank() {
        cudaSetDevice(device);
        for (void* p : opened)
            cudaIpcCloseMemHandle(p);
        cudaFree(d_rows), cudaFree(d_adj0), cudaFree(d_upper_ref), cudaFree(d_upper_adj), cudaFree(d_keys);
        cudaFree(slab), cudaFree(d_vis), cudaFree(d_touched), cudaFree(d_counters), cudaFree(d_in);
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
    }

    PtrBlob export_ptr(void* p) const {
        PtrBlob b;
        memset(&b, 0, sizeof(b));
        b.raw = p;
        if (!local && p)
            LB_CUDA(cudaIpcGetMemHandle(&b.ipc, p));
        return b;
    }
    // a pointer of rank `src` usable on this rank's device; `keep` = stays mapped for the life of the group
    void* import_ptr(const PtrBlob& b, int src, bool keep) {
        if (!b.raw)
            return nullptr;
        if (local || src == rank)
            return b.raw;
        void* p = nullptr;
        LB_CUDA(cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
        if (keep)
            opened.push_back(p);
        return p;
    }
    void release_ptr(void* p, int src) {
        if (!local && src != rank && p)
            LB_CUDA(cudaIpcCloseMemHandle(p));
    }

    // ---- distribute: phase A (root describes its index), B (everybody copies its share), C (map the peers' slabs) ----
    GroupConfigBlob describe(Index* idx) {
        GroupConfigBlob c;
        memset(&c, 0, sizeof(c));
        if (!idx)
            return c;
        idx->flush_staged();
        std::lock_guard<std::mutex> g(idx->mu_);
        if (idx->pending_n_)
            build_pending(*idx);
        if (idx->cfg_.pq)
            throw CudaError("group: pq indexes are not supported");
        if (idx->n_ == 0)
            throw CudaError("group: the index is empty");
        LB_CUDA(cudaDeviceSynchronize());
        c.valid = 1;
        c.metric_kind = idx->cfg_.metric_kind, c.scalar_kind = idx->cfg_.scalar_kind;
        c.dims = idx->cfg_.dims, c.M = idx->cfg_.M, c.M0 = idx->cfg_.M0, c.ef = idx->cfg_.ef;
        c.n = idx->n_, c.upper_lists = idx->upper_lists_, c.row_bytes = idx->row_bytes_, c.vec_bytes = idx->vec_bytes_;
        c.entry = idx->entry_, c.max_level = idx->max_level_;
        c.flags = idx->view().flags;
        { // contiguous row ranges of equal expected work
            unsigned long long* d_hist = nullptr;
            LB_CUDA(cudaMalloc(&d_hist, kBalanceBuckets * sizeof(unsigned long long)));
            LB_CUDA(cudaMemset(d_hist, 0, kBalanceBuckets * sizeof(unsigned long long)));
            indegree_hist_kernel<<<1024, 256>>>(idx->d_adj0_, (size_t)idx->n_ * idx->cfg_.M0, (uint32_t)idx->n_, d_hist);
            LB_CUDA(cudaGetLastError());
            count_launch();
            std::vector<unsigned long long> hist(kBalanceBuckets);
            LB_CUDA(cudaMemcpy(hist.data(), d_hist, hist.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            LB_CUDA(cudaFree(d_hist));
            // weight of a bucket = its in-degree + the mean in-degree: measured on the bench corpus, equal-length ranges leave the
            // first of two ranks 58 % of the evaluations, pure in-degree weighting 45 %; the blend lands near 50 %
            double links = 0;
            for (unsigned long long h : hist)
                links += (double)h;
            const double mean = links / kBalanceBuckets + 1.0;
            double total = 0;
            for (unsigned long long h : hist)
                total += (double)h + mean;
            const size_t n_rows = idx->n_;
            c.bounds[0] = 0;
            double acc = 0;
            int r = 1;
            for (uint32_t b = 0; b < kBalanceBuckets && r < world; ++b) {
                acc += (double)hist[b] + mean;
                while (r < world && acc >= total * r / world) {
                    c.bounds[r] = (uint32_t)std::min<size_t>(n_rows, ((size_t)(b + 1) * n_rows + kBalanceBuckets - 1) / kBalanceBuckets);
                    ++r;
                }
            }
            for (; r <= world; ++r)
                c.bounds[r] = (uint32_t)n_rows;
            c.bounds[world] = (uint32_t)n_rows;
            if (getenv("LB200_GROUP_EQUAL_RANGES")) // experiments: plain n*r/G ranges
                for (int q = 0; q <= world; ++q)
                    c.bounds[q] = (uint32_t)((n_rows * (size_t)q) / (size_t)world);
            for (int q = 1; q <= world; ++q) // monotone, whatever the histogram looked like
                c.bounds[q] = std::max(c.bounds[q], c.bounds[q - 1]);
        }
        c.vectors = export_ptr(idx->d_vectors_), c.adj0 = export_ptr(idx->d_adj0_);
        c.upper_ref = export_ptr(idx->d_upper_ref_), c.upper_adj = export_ptr(idx->d_upper_adj_);
        c.keys = export_ptr(idx->d_keys_);
        return c;
    }

    PtrBlob adopt(const GroupConfigBlob& c, int root, size_t max_batch_, size_t max_results) {
        if (!c.valid)
            throw CudaError("group: the root rank passed no index");
        LB_CUDA(cudaSetDevice(device));
        cfg = IndexConfig();
        cfg.metric_kind = c.metric_kind, cfg.scalar_kind = c.scalar_kind;
        cfg.dims = c.dims, cfg.M = c.M, cfg.M0 = c.M0, cfg.ef = c.ef;
        dist_mode = distance_mode(cfg.metric_kind, cfg.scalar_kind);
        n = c.n, row_bytes = c.row_bytes, vec_bytes = c.vec_bytes, upper_lists = c.upper_lists;
        entry = c.entry, max_level = c.max_level, flags = c.flags;
        if (cfg.M0 > 256)
            throw CudaError("group: connectivity above 128 is not supported");
        for (int r = 0; r <= world; ++r)
            bounds[r] = c.bounds[r]; // contiguous row ranges, sized by the root for equal expected work
        const size_t lo = bounds[rank], hi = bounds[rank + 1];
        LB_CUDA(cudaMalloc(&d_rows, std::max<size_t>((hi - lo) * row_bytes, 16)));
        LB_CUDA(cudaMalloc(&d_adj0, n * cfg.M0 * 4));
        LB_CUDA(cudaMalloc(&d_upper_ref, n * 4));
        LB_CUDA(cudaMalloc(&d_upper_adj, std::max<size_t>(upper_lists * cfg.M * 4, 16)));
        LB_CUDA(cudaMalloc(&d_keys, n * 8));
        auto fetch = [&](void* dst, const PtrBlob& b, size_t offset, size_t bytes) {
            if (!bytes)
                return;
            uint8_t* src = (uint8_t*)import_ptr(b, root, false);
            LB_CUDA(cudaMemcpy(dst, src + offset, bytes, cudaMemcpyDefault));
            release_ptr(src, root);
        };
        fetch(d_rows, c.vectors, lo * row_bytes, (hi - lo) * row_bytes);
        fetch(d_adj0, c.adj0, 0, n * cfg.M0 * 4);
        fetch(d_upper_ref, c.upper_ref, 0, n * 4);
        fetch(d_upper_adj, c.upper_adj, 0, upper_lists * cfg.M * 4);
        fetch(d_keys, c.keys, 0, n * 8);
        // mailboxes
        int sms = 0;
        LB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        cap = (uint32_t)cfg.M0;
        Wmax = (uint32_t)sms * 8u * kGroupWarps; // upper bound of resident warps (8 CTAs per SM)
        max_batch = max_batch_, res_cap = max_results;
        // owner slots: one per query a rank owns in the largest batch (never more than the resident warps)
        Omax = (uint32_t)std::min<size_t>(Wmax, round_up((max_batch + world - 1) / world, kGroupWarps));
        lay = slab_layout(Omax, (uint32_t)world, cap, res_cap, max_batch, row_bytes);
        LB_CUDA(cudaMalloc(&slab, lay.total));
        LB_CUDA(cudaMemset(slab, 0, lay.total));
        // owner scratch: one visited bitmap per owned slot
        words_per_slot = round_up((n + 31) / 32, 32);
        const size_t owned = Omax;
        LB_CUDA(cudaMalloc(&d_vis, owned * words_per_slot * 4));
        LB_CUDA(cudaMemset(d_vis, 0, owned * words_per_slot * 4));
        LB_CUDA(cudaMalloc(&d_touched, owned * (size_t)touched_cap * 4));
        LB_CUDA(cudaMalloc(&d_counters, 16 * sizeof(unsigned long long)));
        LB_CUDA(cudaMemset(d_counters, 0, 16 * sizeof(unsigned long long)));
        // staging of raw input queries for the host-buffer entry point: allocated NOW -- a cudaMalloc between the launches of a
        // search may wait for kernels already running on this device, and those kernels wait for the rank that is allocating
        in_bytes = max_batch * std::max<size_t>(row_bytes, cfg.dims * 4);
        LB_CUDA(cudaMalloc(&d_in, in_bytes));
        LB_CUDA(cudaEventCreate(&ev0));
        LB_CUDA(cudaEventCreate(&ev1));
        LB_CUDA(cudaDeviceSynchronize());
        return export_ptr(slab);
    }

    void map_peers(const PtrBlob* slabs) {
        LB_CUDA(cudaSetDevice(device));
        for (int r = 0; r < world; ++r)
            peer_slab[r] = (uint8_t*)import_ptr(slabs[r], r, true);
    }

    int occupancy_for(uint32_t L) {
        LB_CUDA(cudaSetDevice(device));
        const size_t smem = (size_t)group_warp_layout((uint32_t)row_bytes, group_ring_slots((uint32_t)row_bytes), L, cap).total * kGroupWarps;
        const int nq = pick_nq((uint32_t)row_bytes);
        if (nq < 0)
            throw CudaError("group: vectors wider than 8192 bytes are not supported");
        int occ = 0;
        dispatch_walk(dist_mode, cfg.scalar_kind, nq, [&](auto d, auto s, auto q) {
            occ = group_occupancy_one<decltype(d)::value, decltype(s)::value, decltype(q)::value>(smem);
        });
        int sms = 0;
        LB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        return occ * sms * kGroupWarps;
    }

    uint8_t* qbuf() const { return slab + lay.qbuf; }

    // queries: only the root's are read.  Asynchronous on `stream`.
    void launch(const void* d_queries, size_t nq, size_t stride, int kind, size_t k, uint32_t L, uint32_t W, int root,
                uint64_t* d_out_keys, float* d_out_dists, uint32_t* d_out_counts, cudaStream_t stream) {
        LB_CUDA(cudaSetDevice(device));
        if (nq > max_batch || nq * k > res_cap)
            throw CudaError("group: batch larger than the group was created for (max_batch / max_results)");
        if (nq >= kGroupMaxBatch)
            throw CudaError("group: more than 4 M queries per batch");
        if (rank == root) {
            if (!d_queries)
                throw CudaError("group: the root rank must pass the queries");
            launch_cast_rows(d_queries, stride, kind, qbuf(), row_bytes, cfg.scalar_kind, cfg.dims, nq, stream);
        }
        ++epoch; // callers run renew_flags() first: (epoch & 0xFFF) is never 0 here
        GroupLaunch p;
        memset(&p, 0, sizeof(p));
        p.G = (uint32_t)world, p.me = (uint32_t)rank, p.W = W, p.cap = cap, p.ring_slots = group_ring_slots((uint32_t)row_bytes);
        p.nq = (uint32_t)nq, p.k = (uint32_t)k, p.L = L;
        p.flag_base = (epoch & 0xFFFu) << 20;
        p.epoch = epoch, p.root = (uint32_t)root;
        p.timeout_ns = group_timeout_ns();
        p.g.vectors = d_rows, p.g.adj0 = d_adj0, p.g.upper_ref = d_upper_ref, p.g.upper_adj = d_upper_adj, p.g.keys = d_keys;
        p.g.n = (uint32_t)n, p.g.row_bytes = (uint32_t)row_bytes, p.g.M = (uint32_t)cfg.M, p.g.M0 = (uint32_t)cfg.M0;
        p.g.entry = entry, p.g.max_level = max_level, p.g.flags = flags, p.g.dims = (uint32_t)cfg.dims;
        for (int r = 0; r <= world; ++r)
            p.bounds[r] = bounds[r];
        for (int r = world + 1; r <= kGroupMax; ++r)
            p.bounds[r] = 0xFFFFFFFFu;
        for (int r = 0; r < world; ++r) {
            uint8_t* s = peer_slab[r];
            p.req[r] = (unsigned long long*)(s + lay.req), p.resp[r] = (unsigned long long*)(s + lay.resp);
            p.reqq[r] = s + lay.reqq;
            p.res_keys[r] = (uint64_t*)(s + lay.res_keys), p.res_dists[r] = (float*)(s + lay.res_dists);
            p.res_counts[r] = (uint32_t*)(s + lay.res_counts);
            p.done[r] = (unsigned long long*)(s + lay.done), p.qready[r] = (unsigned long long*)(s + lay.qready);
            p.err[r] = (uint32_t*)(s + lay.err);
        }
        p.queries = peer_slab[root] + lay.qbuf, p.query_stride = (uint32_t)row_bytes;
        p.vis = d_vis, p.touched = d_touched, p.words_per_slot = words_per_slot, p.touched_cap = touched_cap;
        p.counters = d_counters;
        LB_CUDA(cudaMemsetAsync(d_counters, 0, 16 * sizeof(unsigned long long), stream));
        const size_t smem = (size_t)group_warp_layout((uint32_t)row_bytes, group_ring_slots((uint32_t)row_bytes), L, cap).total * kGroupWarps;
        uint32_t O = 0, H = 0;
        if (!group_plan(world, nq, W, Omax, O, H))
            throw CudaError("group: not enough resident warps for this batch's mailboxes");
        p.O = O, p.H = H, p.W = W;
        const uint32_t grid = (O + H) / kGroupWarps;
        const int nqc = pick_nq((uint32_t)row_bytes);
        LB_CUDA(cudaEventRecord(ev0, stream));
        dispatch_walk(dist_mode, cfg.scalar_kind, nqc, [&](auto d, auto s, auto q) {
            group_launch_one<decltype(d)::value, decltype(s)::value, decltype(q)::value>(p, grid, smem, stream);
        });
        LB_CUDA(cudaEventRecord(ev1, stream));
        if (d_out_keys || d_out_dists || d_out_counts) {
            const size_t total = std::max(nq * k, nq);
            group_copy_results_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
                (const uint64_t*)(slab + lay.res_keys), (const float*)(slab + lay.res_dists), (const uint32_t*)(slab + lay.res_counts),
                d_out_keys, d_out_dists, d_out_counts, nq, k);
            LB_CUDA(cudaGetLastError());
            count_launch();
        }
        last_nq = (uint32_t)nq;
    }

    void check_error() {
        uint32_t e = 0;
        LB_CUDA(cudaMemcpy(&e, slab + lay.err, 4, cudaMemcpyDeviceToHost));
        if (e)
            throw CudaError("group: a rank timed out waiting for a peer (rank " + std::to_string(e - 1) + " gave up first)");
    }
};

// ---- the handle behind lb200_group_t: one rank of a multi-process group, or all ranks of a single-process one --------
struct Group {
    std::vector<GroupRank*> ranks; // 1 entry (multi-process) or `world` entries (single process)
    int world = 1, root = 0;
    lb200_allgather_fn ag = nullptr;
    void* ag_ctx = nullptr;
    bool distributed = false;
    std::mutex mu;
    std::vector<cudaStream_t> streams; // single-process: one per device
    ~Group() {
        for (GroupRank* r : ranks)
            delete r;
        for (size_t i = 0; i < streams.size(); ++i)
            if (streams[i])
                cudaStreamDestroy(streams[i]);
    }
    bool local() const { return ranks.size() > 1 || world == 1; }

    // every rank contributes `bytes`; everybody receives world * bytes
    void exchange(const std::vector<std::vector<uint8_t>>& mine, std::vector<uint8_t>& all, size_t bytes) {
        all.assign((size_t)world * bytes, 0);
        if (ranks.size() == (size_t)world) {
            for (int r = 0; r < world; ++r)
                memcpy(all.data() + (size_t)r * bytes, mine[r].data(), bytes);
        } else {
            ag(ag_ctx, mine[0].data(), all.data(), bytes);
        }
    }
};

static uint32_t agree_W(Group& G, uint32_t L) {
    GroupRank* r0 = G.ranks[0];
    auto it = r0->W_for_L.find(L);
    if (it != r0->W_for_L.end())
        return it->second;
    // every rank reports how many warps its device keeps resident for this beam width, and which physical device it is:
    // ranks that share a device (a 1-GPU box running the G-rank protocol) must be resident together and split its warps
    struct Offer {
        uint32_t w;
        char uuid[16];
    };
    std::vector<std::vector<uint8_t>> mine(G.ranks.size(), std::vector<uint8_t>(sizeof(Offer)));
    for (size_t i = 0; i < G.ranks.size(); ++i) {
        Offer o;
        o.w = (uint32_t)G.ranks[i]->occupancy_for(L);
        cudaDeviceProp prop;
        LB_CUDA(cudaGetDeviceProperties(&prop, G.ranks[i]->device));
        memcpy(o.uuid, prop.uuid.bytes, 16);
        memcpy(mine[i].data(), &o, sizeof(o));
    }
    std::vector<uint8_t> all;
    G.exchange(mine, all, sizeof(Offer));
    const Offer* offers = (const Offer*)all.data();
    uint32_t W = 0xFFFFFFFFu, share = 1;
    for (int r = 0; r < G.world; ++r) {
        W = std::min(W, offers[r].w);
        uint32_t same = 0;
        for (int t = 0; t < G.world; ++t)
            same += memcmp(offers[r].uuid, offers[t].uuid, 16) == 0;
        share = std::max(share, same);
    }
    W = std::min(W, r0->Wmax);
    W /= share;
    W -= W % (uint32_t)(G.world * kGroupWarps); // whole CTAs, and a slot keeps its owner across waves
    if (W == 0)
        throw CudaError("group: the search kernel does not fit on an SM (ef too large for shared memory)");
    for (GroupRank* r : G.ranks)
        r->W_for_L[L] = W;
    return W;
}

void group_distribute(Group& G, Index* root_index, int root, size_t max_batch, size_t max_results) {
    DeviceGuard dg;
    std::lock_guard<std::mutex> lk(G.mu);
    if (G.distributed)
        throw CudaError("group: already distributed");
    if (root < 0 || root >= G.world)
        throw CudaError("group: bad root rank");
    const size_t nr = G.ranks.size();
    // A: the root describes its index
    std::vector<std::vector<uint8_t>> mine(nr, std::vector<uint8_t>(sizeof(GroupConfigBlob)));
    for (size_t i = 0; i < nr; ++i) {
        GroupRank* r = G.ranks[i];
        LB_CUDA(cudaSetDevice(r->device));
        GroupConfigBlob c = r->describe(r->rank == root ? root_index : nullptr);
        memcpy(mine[i].data(), &c, sizeof(c));
    }
    std::vector<uint8_t> all;
    G.exchange(mine, all, sizeof(GroupConfigBlob));
    GroupConfigBlob cfg;
    memcpy(&cfg, all.data() + (size_t)root * sizeof(GroupConfigBlob), sizeof(cfg));
    // B: every rank copies its row range and the graph from the root, allocates its mailboxes
    std::vector<std::vector<uint8_t>> slabs(nr, std::vector<uint8_t>(sizeof(PtrBlob)));
    for (size_t i = 0; i < nr; ++i) {
        PtrBlob b = G.ranks[i]->adopt(cfg, root, max_batch, max_results);
        memcpy(slabs[i].data(), &b, sizeof(b));
    }
    G.exchange(slabs, all, sizeof(PtrBlob)); // also tells the root that everybody is done reading its index
    // C: map the peers' mailboxes
    for (size_t i = 0; i < nr; ++i)
        G.ranks[i]->map_peers((const PtrBlob*)all.data());
    std::vector<std::vector<uint8_t>> tick(nr, std::vector<uint8_t>(4, 0));
    G.exchange(tick, all, 4); // nobody searches before everybody has mapped everybody
    G.root = root;
    G.distributed = true;
}

// Message flags are (epoch mod 4096) << 20 | sequence number.  Before the 12-bit part wraps, every rank waits for its last
// launch (which ends only when all ranks have finished theirs), all ranks meet, clear their mailboxes and meet again:
// a word written 4096 launches ago can then never be mistaken for a fresh one.  Collective; costs two exchanges per 4095
// searches.
static void renew_flags(Group& G, const std::vector<cudaStream_t>& streams) {
    GroupRank* r0 = G.ranks[0];
    if (((r0->epoch + 1) & 0xFFFu) != 0u)
        return;
    std::vector<std::vector<uint8_t>> tick(G.ranks.size(), std::vector<uint8_t>(4, 0));
    std::vector<uint8_t> all;
    for (size_t i = 0; i < G.ranks.size(); ++i) {
        LB_CUDA(cudaSetDevice(G.ranks[i]->device));
        LB_CUDA(cudaStreamSynchronize(streams[i]));
    }
    G.exchange(tick, all, 4);
    for (GroupRank* r : G.ranks) {
        LB_CUDA(cudaSetDevice(r->device));
        LB_CUDA(cudaMemset(r->slab + r->lay.req, 0, r->lay.reqq - r->lay.req)); // request + response mailboxes
        LB_CUDA(cudaDeviceSynchronize());
        r->epoch += 1; // skip the value whose 12-bit part is 0 (flag 0 means "never written")
    }
    G.exchange(tick, all, 4);
}

static uint32_t beam_width(const GroupRank& r, size_t k, size_t ef) {
    size_t L = ef ? ef : r.cfg.ef;
    if (L < k)
        L = k; // index.hpp:2706
    if (L > 4096)
        throw CudaError("search: max(ef, count) > 4096 is not supported");
    return (uint32_t)L;
}

void group_search_device(Group& G, const void* d_queries, size_t nq, size_t stride, int kind, size_t k, size_t ef,
                         uint64_t* d_keys, float* d_dists, uint32_t* d_counts, cudaStream_t stream) {
    DeviceGuard dg;
    std::lock_guard<std::mutex> lk(G.mu);
    if (!G.distributed)
        throw CudaError("group: lb200_group_distribute has not been called");
    if (G.ranks.size() != 1)
        throw CudaError("group: device-buffer search is the multi-process entry point; use lb200_group_search_batch");
    if (!nq || !k)
        return;
    GroupRank* r = G.ranks[0];
    const uint32_t L = beam_width(*r, k, ef);
    const uint32_t W = agree_W(G, L);
    renew_flags(G, std::vector<cudaStream_t>(1, stream));
    r->launch(d_queries, nq, stride, kind, k, L, W, G.root, d_keys, d_dists, d_counts, stream);
}

// host buffers: queries are read on the root rank only; every rank receives the results
void group_search_host(Group& G, const void* queries, size_t nq, size_t stride, int kind, size_t k, size_t ef, uint64_t* keys,
                       float* dists, size_t* counts) {
    DeviceGuard dg;
    std::lock_guard<std::mutex> lk(G.mu);
    if (!G.distributed)
        throw CudaError("group: lb200_group_distribute has not been called");
    if (!nq || !k)
        return;
    const bool trace = getenv("LB200_GROUP_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0;
    const size_t nr = G.ranks.size();
    GroupRank* r0 = G.ranks[0];
    const uint32_t L = beam_width(*r0, k, ef);
    const uint32_t W = agree_W(G, L);
    const size_t in_bytes = scalar_row_bytes(kind, r0->cfg.dims);
    if (G.streams.size() != nr) {
        G.streams.assign(nr, nullptr);
        for (size_t i = 0; i < nr; ++i) {
            LB_CUDA(cudaSetDevice(G.ranks[i]->device));
            LB_CUDA(cudaStreamCreateWithFlags(&G.streams[i], cudaStreamNonBlocking));
        }
    }
    renew_flags(G, G.streams);
    size_t first = 0; // the root goes first: its (possibly pageable, hence host-blocking) upload must not sit behind kernels that
    for (size_t i = 0; i < nr; ++i) // already spin on the same device waiting for the root's queries
        if (G.ranks[i]->rank == G.root)
            first = i;
    for (size_t step = 0; step < nr; ++step) {
        const size_t i = (first + step) % nr;
        GroupRank* r = G.ranks[i];
        LB_CUDA(cudaSetDevice(r->device));
        const void* dq = nullptr;
        if (r->rank == G.root) {
            if (!queries)
                throw CudaError("group: the root rank must pass the queries");
            if (nq * in_bytes > r->in_bytes) // (sized for max_batch at distribution time; nothing may be allocated here)
                throw CudaError("group: batch larger than the group was created for (max_batch)");
            LB_CUDA(cudaMemcpy2DAsync(r->d_in, in_bytes, queries, stride, in_bytes, nq, cudaMemcpyHostToDevice, G.streams[i]));
            dq = r->d_in;
        }
        r->launch(dq, nq, in_bytes, kind, k, L, W, G.root, nullptr, nullptr, nullptr, G.streams[i]);
    }
    const double t1 = trace ? now() : 0;
    // results: from the first rank of this process (every rank holds all of them); the error word rides along
    LB_CUDA(cudaSetDevice(r0->device));
    cudaStream_t s0 = G.streams[0];
    if (keys)
        LB_CUDA(cudaMemcpyAsync(keys, r0->slab + r0->lay.res_keys, nq * k * 8, cudaMemcpyDeviceToHost, s0));
    if (dists)
        LB_CUDA(cudaMemcpyAsync(dists, r0->slab + r0->lay.res_dists, nq * k * 4, cudaMemcpyDeviceToHost, s0));
    if (counts) {
        r0->h_counts.resize(nq);
        LB_CUDA(cudaMemcpyAsync(r0->h_counts.data(), r0->slab + r0->lay.res_counts, nq * 4, cudaMemcpyDeviceToHost, s0));
    }
    for (size_t i = 0; i < nr; ++i) {
        GroupRank* r = G.ranks[i];
        LB_CUDA(cudaSetDevice(r->device));
        LB_CUDA(cudaMemcpyAsync(&r->h_err, r->slab + r->lay.err, 4, cudaMemcpyDeviceToHost, G.streams[i]));
    }
    for (size_t i = 0; i < nr; ++i) {
        LB_CUDA(cudaSetDevice(G.ranks[i]->device));
        LB_CUDA(cudaStreamSynchronize(G.streams[i]));
    }
    const double t2 = trace ? now() : 0;
    for (size_t i = 0; i < nr; ++i)
        if (G.ranks[i]->h_err)
            throw CudaError("group: a rank timed out waiting for a peer (rank " + std::to_string(G.ranks[i]->h_err - 1) + " gave up first)");
    if (counts)
        for (size_t i = 0; i < nq; ++i)
            counts[i] = r0->h_counts[i];
    if (trace)
        fprintf(stderr, "lb200 group rank %d: issue %.3f ms, wait %.3f ms\n", r0->rank, t1 - t0, t2 - t1);
}

void group_stats(Group& G, int which, GroupStats& out) {
    DeviceGuard dg;
    std::lock_guard<std::mutex> lk(G.mu);
    if (which < 0 || (size_t)which >= G.ranks.size())
        throw CudaError("group: no such local rank");
    GroupRank* r = G.ranks[which];
    LB_CUDA(cudaSetDevice(r->device));
    LB_CUDA(cudaDeviceSynchronize());
    r->check_error();
    unsigned long long c[16] = {0};
    LB_CUDA(cudaMemcpy(c, r->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    memset(&out, 0, sizeof(out));
    out.rank = r->rank, out.world = r->world;
    out.queries = r->last_nq;
    out.owner_computed_distances = c[1], out.owner_base_pops = c[2], out.owner_upper_hops = c[3], out.owner_rounds = c[4];
    out.local_rows_evaluated = c[5];
    out.owner_cycles_produce = c[8], out.owner_cycles_local = c[9], out.owner_cycles_wait = c[10], out.owner_cycles_consume = c[11];
    out.local_row_bytes = c[5] * r->vec_bytes;
    out.rows_held = r->bounds[r->rank + 1] - r->bounds[r->rank];
    float ms = 0.f;
    if (r->last_nq && cudaEventElapsedTime(&ms, r->ev0, r->ev1) == cudaSuccess)
        out.kernel_ms = ms;
    else
        (void)cudaGetLastError();
}

Group* group_create_ipc(int rank, int world, lb200_allgather_fn ag, void* ctx) {
    if (world < 1 || world > kGroupMax || rank < 0 || rank >= world)
        throw CudaError("group: world size must be 1..8 and 0 <= rank < world");
    if (world > 1 && !ag)
        throw CudaError("group: an all-gather callback is required to bootstrap a multi-process group");
    require_device();
    Group* G = new Group();
    G->world = world, G->ag = ag, G->ag_ctx = ctx;
    GroupRank* r = new GroupRank();
    r->rank = rank, r->world = world, r->local = (world == 1);
    LB_CUDA(cudaGetDevice(&r->device));
    G->ranks.push_back(r);
    return G;
}

Group* group_create_local(const int* devices, int ndev) {
    if (ndev < 1 || ndev > kGroupMax)
        throw CudaError("group: 1..8 devices");
    require_device();
    DeviceGuard dg;
    int have = 0;
    LB_CUDA(cudaGetDeviceCount(&have));
    Group* G = new Group();
    G->world = ndev;
    try {
        for (int i = 0; i < ndev; ++i) {
            const int dev = devices ? devices[i] : i;
            if (dev < 0 || dev >= have)
                throw CudaError("group: no such CUDA device");
            GroupRank* r = new GroupRank();
            r->rank = i, r->world = ndev, r->local = true, r->device = dev;
            G->ranks.push_back(r);
        }
        for (int i = 0; i < ndev; ++i) {
            LB_CUDA(cudaSetDevice(G->ranks[i]->device));
            for (int j = 0; j < ndev; ++j) {
                if (i == j || G->ranks[i]->device == G->ranks[j]->device)
                    continue;
                int can = 0;
                LB_CUDA(cudaDeviceCanAccessPeer(&can, G->ranks[i]->device, G->ranks[j]->device));
                if (!can)
                    throw CudaError("group: the devices cannot access each other's memory (no NVLink / P2P)");
                cudaError_t e = cudaDeviceEnablePeerAccess(G->ranks[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    LB_CUDA(e);
                (void)cudaGetLastError();
            }
        }
    } catch (...) {
        delete G;
        throw;
    }
    return G;
}

// ---- the same kernel as a single-GPU search path: G = 1, every slot is an owner, nothing crosses NVLink ------------------
// One warp per query instead of one CTA per query (search.cu): no CTA-wide barriers, 28 queries in flight per SM instead of
// 8-10.  For narrow rows (binary vectors, short f16/i8 rows) the per-expansion serial chain of the CTA kernel -- warp 0
// working while three warps wait at the barrier -- is what bounds it; here every warp always has its own work.
void launch_warp_search(Index& idx, const uint8_t* qbuf, size_t qrow, size_t nq, size_t k, uint32_t L, uint64_t* d_keys, float* d_dists,
                        uint32_t* d_counts, cudaStream_t stream) {
    const IndexConfig& cfg = idx.cfg_;
    if (cfg.pq)
        throw CudaError("warp search: pq indexes use the CTA kernel");
    const uint32_t cap = (uint32_t)cfg.M0;
    if (cap > 256)
        throw CudaError("warp search: connectivity above 128 is not supported");
    const int nqc = pick_nq((uint32_t)idx.row_bytes_);
    const size_t smem = (size_t)group_warp_layout((uint32_t)idx.row_bytes_, group_ring_slots((uint32_t)idx.row_bytes_), L, cap).total * kGroupWarps;
    int occ = 0;
    dispatch_walk(idx.dist_mode_, cfg.scalar_kind, nqc, [&](auto d, auto s2, auto q) {
        occ = group_occupancy_one<decltype(d)::value, decltype(s2)::value, decltype(q)::value>(smem);
    });
    if (occ < 1)
        throw CudaError("warp search: kernel does not fit on an SM (ef too large for shared memory)");
    const uint32_t W = (uint32_t)occ * (uint32_t)device_sm_count() * kGroupWarps;
    idx.ensure_scratch(W); // one visited bitmap + touched list per slot
    const size_t aux_need = 256 + round_up(nq * 4, 256);
    if (aux_need > idx.warp_aux_bytes_) {
        if (idx.d_warp_aux_)
            LB_CUDA(cudaFree(idx.d_warp_aux_));
        idx.d_warp_aux_ = nullptr;
        idx.warp_aux_bytes_ = aux_need + aux_need / 2;
        LB_CUDA(cudaMalloc(&idx.d_warp_aux_, idx.warp_aux_bytes_));
    }
    uint8_t* aux = idx.d_warp_aux_;
    LB_CUDA(cudaMemsetAsync(aux, 0, 256, stream)); // [0,128) counters, [128,136) done, [136,144) qready, [144,148) err
    GroupLaunch p;
    memset(&p, 0, sizeof(p));
    p.G = 1, p.me = 0, p.W = W, p.cap = cap, p.ring_slots = group_ring_slots((uint32_t)idx.row_bytes_);
    p.nq = (uint32_t)nq, p.k = (uint32_t)k, p.L = L;
    p.flag_base = 1u << 20, p.epoch = 1, p.root = 0;
    p.timeout_ns = ~0ull;
    p.g = idx.view();
    p.bounds[0] = 0;
    for (int r = 1; r <= kGroupMax; ++r)
        p.bounds[r] = 0xFFFFFFFFu;
    p.res_keys[0] = d_keys, p.res_dists[0] = d_dists;
    p.res_counts[0] = d_counts ? d_counts : (uint32_t*)(aux + 256);
    p.counters = (unsigned long long*)aux;
    p.done[0] = (unsigned long long*)(aux + 128), p.qready[0] = (unsigned long long*)(aux + 136), p.err[0] = (uint32_t*)(aux + 144);
    p.queries = qbuf, p.query_stride = (uint32_t)qrow;
    p.vis = idx.scratch_.visited, p.touched = idx.scratch_.touched;
    p.words_per_slot = idx.scratch_.words_per_cta, p.touched_cap = idx.scratch_.touched_cap;
    p.O = (uint32_t)std::min<size_t>(W & ~3u, round_up(nq, kGroupWarps)), p.H = 0;
    const uint32_t grid = p.O / kGroupWarps;
    dispatch_walk(idx.dist_mode_, cfg.scalar_kind, nqc, [&](auto d, auto s2, auto q) {
        group_launch_one<decltype(d)::value, decltype(s2)::value, decltype(q)::value>(p, grid, smem, stream);
    });
    // work counters where Index::last_stats() reads them: [1] distance evaluations, [2] base pops, [3] upper hops
    LB_CUDA(cudaMemcpyAsync(idx.scratch_.counters + 1, p.counters + 1, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    LB_CUDA(cudaMemcpyAsync(idx.scratch_.counters + 4, p.counters + 7, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
}

void group_free(Group* G) {
    DeviceGuard dg;
    delete G;
}
int group_world(const Group& G) { return G.world; }
int group_local_ranks(const Group& G) { return (int)G.ranks.size(); }

} // namespace lb200
