// lantern_b200 -- exhaustive search on the 5th-generation tensor cores (tcgen05 + TMEM), f32 l2sq / cos.
//
// The GEMM-shaped piece of the reference path: exact_search_t (U/include/usearch/index_plugins.hpp:1582-1675) behind
// usearch_exact_search (U/c/lib.cpp:450-481) measures every query against every row.  A tensor core cannot reproduce
// the reference's fp32 sum of (a-b)^2 bit for bit, so it is used as a FILTER with an error bound, and the answer is
// produced by the same fp32 arithmetic as before:
//   1. norms_kernel            |x|^2 of every row and query (fp32)
//   2. exact_tc_filter_kernel  D = Q . X^T on tcgen05 in "3xTF32" (a = hi + lo with hi, lo representable in tf32;
//                              q.x ~ qhi.xhi + qhi.xlo + qlo.xhi, fp32 accumulation in TMEM), fused epilogue: distance
//                              from norms and dot product, per-query candidate list (a max-heap) of KP > k entries per row
//                              group, plus the smallest lower bound of anything the list had to drop
//   3. exact_tc_rerank_kernel  per query: candidates whose lower bound (approx - eps) does not exceed the k-th smallest
//                              upper bound (approx + eps) are re-measured with the SIMT distance code of exact.cu
//                              (same lane layout, same reduction) and sorted by (distance, offset): ids and distances are
//                              those of exact_block_kernel, ties included.  eps = (dims + 32) * 2^-24 * |q||x| covers the
//                              split's truncation and worst-case fp32 accumulation on both sides.  A query whose lists
//                              may have dropped a qualifying row, or with more qualifying rows than fit, is flagged and
//                              answered by the SIMT kernels (exact.cu) -- never silently wrong.
//
// Filter kernel anatomy (one CTA per SM, 9 warps, persistent over its row tiles):
//   warps 0-3  epilogue   tcgen05.ld 32 lanes x 32 columns per instruction; thread t owns query t of the tile
//   warps 4-7  producers  rows come in as fp32 (LDG.128, 8 consecutive chunks of a row per thread), are split into hi / lo
//                         in registers and stored (STS.128) straight into the K-major no-swizzle core-matrix layout the
//                         UMMA descriptors describe: address(row, chunk) = (row / 8) * 1024 + chunk * 128 + (row % 8) * 16
//                         per 32-float k-block (SBO = 1024 B, LBO = 128 B)
//   warp 8     MMA        one elected thread: per k-block 4 x 3 tcgen05.mma.kind::tf32 (M = 128 queries, N = 256 rows,
//                         K = 8), tcgen05.commit to the stage's "empty" barrier; per tile one commit to "tmem full"
// Two shared-memory stages (96 KB each: A hi/lo 2 x 16 KB, B hi/lo 2 x 32 KB) and two TMEM accumulators (2 x 256 columns)
// keep the loads of tile t+1 and the epilogue of tile t-1 under the MMAs of tile t.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <type_traits>

#include "distance.cuh"
#include "engine.h"

namespace lb200 {

namespace {

constexpr int kTcThreads = 288;          // 4 epilogue + 4 producer + 1 MMA warp
constexpr uint32_t kTileM = 128;         // queries per tile (UMMA M)
constexpr uint32_t kTileN = 256;         // rows per tile (UMMA N)
constexpr uint32_t kBlockK = 32;         // floats per k-block (128 B of a row)
constexpr uint32_t kStages = 2;
constexpr uint32_t kABytes = kTileM * kBlockK * 4; // 16 KB
constexpr uint32_t kBBytes = kTileN * kBlockK * 4; // 32 KB
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
constexpr uint32_t kTcSmem = kStages * kStageBytes + 1024;
constexpr uint32_t kMaxKP = 256;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                 "}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous; LBO = distance between core matrices adjacent in K,
// SBO = distance between 8-row groups (both in 16-byte units in the descriptor); version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46; // version_
    return d;               // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// kind::tf32 instruction descriptor: D = f32, A = B = tf32, both K-major, N = 256, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const uint4& v, uint4& hi, uint4& lo) {
    const float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y), x2 = __uint_as_float(v.z), x3 = __uint_as_float(v.w);
    const float h0 = tf32_round(x0), h1 = tf32_round(x1), h2 = tf32_round(x2), h3 = tf32_round(x3);
    hi = make_uint4(__float_as_uint(h0), __float_as_uint(h1), __float_as_uint(h2), __float_as_uint(h3));
    lo = make_uint4(__float_as_uint(tf32_round(x0 - h0)), __float_as_uint(tf32_round(x1 - h1)), __float_as_uint(tf32_round(x2 - h2)),
                    __float_as_uint(tf32_round(x3 - h3)));
}

__global__ void norms_kernel(const uint8_t* __restrict__ rows, size_t n, size_t stride, uint32_t nchunks, float* __restrict__ out) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n)
        return;
    const uint4* p = reinterpret_cast<const uint4*>(rows + row * stride);
    float s = 0.f;
    for (uint32_t c = lane; c < nchunks; c += 32) {
        const uint4 v = __ldg(p + c);
        const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), cc = __uint_as_float(v.z), d = __uint_as_float(v.w);
        s += a * a + b * b + cc * cc + d * d;
    }
    s = warp_sum(s);
    if (lane == 0)
        out[row] = s;
}

struct TcParams {
    const uint8_t* data;
    size_t n, data_stride;
    const uint8_t* queries;
    uint32_t nq;
    size_t q_stride;
    uint32_t nchunks; // 16-byte chunks per row (row_bytes / 16)
    const float* xn;
    const float* qn;
    uint32_t groups;  // row groups per query tile: CTA (qt, g) handles row tiles g, g + groups, ...
    uint32_t row_tiles;
    uint32_t KP;
    float eps_rel;
    float* cand_d;    // [nq][groups][KP] approx distances of the kept rows, heap order (+inf padded)
    uint32_t* cand_i; // [nq][groups][KP]
    float* dropped_lb; // [nq][groups] smallest lower bound among entries the list could not keep (+inf if none)
};

template <int DM> __device__ __forceinline__ float approx_distance(float dot, float qn, float xn) {
    if constexpr (DM == DM_COS) {
        if (qn == 0.f && xn == 0.f)
            return 0.f;
        if (qn == 0.f || xn == 0.f)
            return 1.f;
        return 1.f - dot / (sqrtf(qn) * sqrtf(xn));
    } else {
        return qn + xn - 2.f * dot;
    }
}
template <int DM> __device__ __forceinline__ float approx_eps(float eps_rel, float qn, float xn) {
    if constexpr (DM == DM_COS)
        return eps_rel + 3e-6f; // |dot error| / (|q||x|), plus norms, square roots and the division in fp32
    else
        return 2.f * eps_rel * sqrtf(qn) * sqrtf(xn) + 2e-6f * (qn + xn); // dot error; norms and the final sum in fp32
}

template <int DM> __global__ void __launch_bounds__(kTcThreads, 1) exact_tc_filter_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qt = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const uint32_t q0 = qt * kTileM;
    const uint32_t KB = (p.nchunks * 4 + kBlockK - 1) / kBlockK; // k-blocks per row
    const uint32_t my_tiles = g < p.row_tiles ? (p.row_tiles - g + p.groups - 1) / p.groups : 0;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 4);  // one arrive per producer warp
            mbar_init(&empty_bar[s], 1); // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);  // tcgen05.commit
            mbar_init(&tempty_bar[b], 4); // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 8) { // TMEM: all 512 columns (two 128 x 256 fp32 accumulators); this warp also frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= 4 && warp < 8) {
        // ===== producers: global fp32 -> (hi, lo) tf32 pairs in the UMMA core-matrix layout =====
        const int pw = warp - 4;
        uint32_t it = 0;
        for (uint32_t t = 0; t < my_tiles; ++t) {
            const size_t row0 = (size_t)(g + t * p.groups) * kTileN;
            for (uint32_t kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % kStages;
                uint8_t* st = smem + (size_t)s * kStageBytes;
                // A: 128 query rows = 4 slabs of 32 rows, one per producer warp; B: 256 corpus rows = 8 slabs, two per warp;
                // lane = row within the slab.  All 24 16-byte loads of the k-block are issued before the first conversion
                // (one memory latency per k-block instead of three), and the thread's three 128-byte lines of the NEXT
                // k-block are requested into L2 meanwhile.
                const uint32_t ra = pw * 32 + lane, qa = q0 + ra;
                const uint32_t rb0 = (pw * 2) * 32 + lane, rb1 = rb0 + 32;
                const size_t row_b0 = row0 + rb0, row_b1 = row0 + rb1;
                const uint4* sa = reinterpret_cast<const uint4*>(p.queries + (size_t)qa * p.q_stride) + kb * 8;
                const uint4* sb0 = reinterpret_cast<const uint4*>(p.data + row_b0 * p.data_stride) + kb * 8;
                const uint4* sb1 = reinterpret_cast<const uint4*>(p.data + row_b1 * p.data_stride) + kb * 8;
                uint4 va[8], vb0[8], vb1[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const bool in_k = kb * 8 + c < p.nchunks;
                    va[c] = (qa < p.nq && in_k) ? __ldg(sa + c) : make_uint4(0, 0, 0, 0);
                    vb0[c] = (row_b0 < p.n && in_k) ? __ldg(sb0 + c) : make_uint4(0, 0, 0, 0);
                    vb1[c] = (row_b1 < p.n && in_k) ? __ldg(sb1 + c) : make_uint4(0, 0, 0, 0);
                }
                if ((kb + 1) * 8 < p.nchunks) {
                    if (row_b0 < p.n)
                        prefetch_l2(sb0 + 8);
                    if (row_b1 < p.n)
                        prefetch_l2(sb1 + 8);
                }
                mbar_wait(&empty_bar[s], ((it / kStages) & 1u) ^ 1u); // (the loads above are already in flight)
                uint8_t* da = st + (ra >> 3) * 1024 + (ra & 7) * 16;
                uint8_t* db0 = st + 2 * kABytes + (rb0 >> 3) * 1024 + (rb0 & 7) * 16;
                uint8_t* db1 = st + 2 * kABytes + (rb1 >> 3) * 1024 + (rb1 & 7) * 16;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 hi, lo;
                    split4(va[c], hi, lo);
                    *reinterpret_cast<uint4*>(da + c * 128) = hi;
                    *reinterpret_cast<uint4*>(da + kABytes + c * 128) = lo;
                    split4(vb0[c], hi, lo);
                    *reinterpret_cast<uint4*>(db0 + c * 128) = hi;
                    *reinterpret_cast<uint4*>(db0 + kBBytes + c * 128) = lo;
                    split4(vb1[c], hi, lo);
                    *reinterpret_cast<uint4*>(db1 + c * 128) = hi;
                    *reinterpret_cast<uint4*>(db1 + kBBytes + c * 128) = lo;
                }
                fence_proxy_async(); // generic-proxy stores above -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&full_bar[s]);
            }
        }
    } else if (warp == 8) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTileM, kTileN);
            uint32_t it = 0;
            for (uint32_t t = 0; t < my_tiles; ++t) {
                const uint32_t b = t & 1u;
                mbar_wait(&tempty_bar[b], ((t >> 1) & 1u) ^ 1u); // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + b * kTileN;
                for (uint32_t kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % kStages;
                    mbar_wait(&full_bar[s], (it / kStages) & 1u);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes), a_lo = a_hi + kABytes;
                    const uint32_t b_hi = a_hi + 2 * kABytes, b_lo = b_hi + kBBytes;
#pragma unroll
                    for (uint32_t kk = 0; kk < kBlockK / 8; ++kk) { // UMMA K = 8 tf32 = two 16-byte core-matrix columns
                        const uint32_t off = kk * 256;
                        const uint64_t dah = umma_desc(a_hi + off, 128, 1024), dal = umma_desc(a_lo + off, 128, 1024);
                        const uint64_t dbh = umma_desc(b_hi + off, 128, 1024), dbl = umma_desc(b_lo + off, 128, 1024);
                        tc_mma_tf32(d_tmem, dah, dbh, idesc, (kb | kk) ? 1u : 0u);
                        tc_mma_tf32(d_tmem, dah, dbl, idesc, 1u);
                        tc_mma_tf32(d_tmem, dal, dbh, idesc, 1u);
                    }
                    tc_commit(&empty_bar[s]); // the stage may be refilled once these MMAs have read it
                }
                tc_commit(&tfull_bar[b]); // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: thread (warp, lane) owns query q0 + 32 * warp + lane == TMEM lane 32 * warp + lane =====
        const uint32_t q = q0 + warp * 32 + lane;
        const bool live = q < p.nq;
        const float qn = live ? __ldg(p.qn + q) : 0.f;
        float ld[kMaxKP];
        uint32_t li[kMaxKP];
        uint32_t size = 0;
        const uint32_t KP = p.KP;
        float thr = INFINITY, dropped = INFINITY;
        for (uint32_t t = 0; t < my_tiles; ++t) {
            const uint32_t b = t & 1u;
            const size_t row0 = (size_t)(g + t * p.groups) * kTileN;
            mbar_wait(&tfull_bar[b], (t >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < kTileN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + b * kTileN + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                               "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                               "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr)
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const size_t row = row0 + c0 + j;
                        if (row < p.n) {
                            const float xn = __ldg(p.xn + row);
                            const float d = approx_distance<DM>(__uint_as_float(r[j]), qn, xn);
                            if (d < thr) {
                                // max-heap on the distance (root = the list's current worst, = thr once the list is full):
                                // log2(KP) steps per accepted row.  A sorted array cost KP/2 shifts per insert, and every
                                // insert of ANY lane stalls the whole warp (ncu: the shift loop was the kernel's hot spot).
                                uint32_t pos;
                                if (size < KP) {
                                    pos = size++;
                                    while (pos > 0) {
                                        const uint32_t par = (pos - 1) >> 1;
                                        if (ld[par] >= d)
                                            break;
                                        ld[pos] = ld[par], li[pos] = li[par];
                                        pos = par;
                                    }
                                } else { // replace the root; the evicted row is remembered through its lower bound
                                    dropped = fminf(dropped, ld[0] - approx_eps<DM>(p.eps_rel, qn, __ldg(p.xn + li[0])));
                                    pos = 0;
                                    for (;;) {
                                        uint32_t c = 2 * pos + 1;
                                        if (c >= KP)
                                            break;
                                        if (c + 1 < KP && ld[c + 1] > ld[c])
                                            ++c;
                                        if (ld[c] <= d)
                                            break;
                                        ld[pos] = ld[c], li[pos] = li[c];
                                        pos = c;
                                    }
                                }
                                ld[pos] = d, li[pos] = (uint32_t)row;
                                if (size == KP)
                                    thr = ld[0];
                            } else {
                                dropped = fminf(dropped, d - approx_eps<DM>(p.eps_rel, qn, xn));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&tempty_bar[b]);
        }
        if (live) {
            float* od = p.cand_d + ((size_t)q * p.groups + g) * KP;
            uint32_t* oi = p.cand_i + ((size_t)q * p.groups + g) * KP;
            for (uint32_t i = 0; i < KP; ++i) {
                od[i] = i < size ? ld[i] : INFINITY;
                oi[i] = i < size ? li[i] : 0xFFFFFFFFu;
            }
            p.dropped_lb[(size_t)q * p.groups + g] = dropped;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- re-rank: exact fp32 distances (exact.cu's arithmetic) of the candidates that can still be in the top k ----------
constexpr uint32_t kRrThreads = 256;
constexpr uint32_t kRrCap = 512; // candidates that survive the bound test, per query

struct RrParams {
    const uint8_t* data;
    size_t n, data_stride;
    const uint8_t* queries;
    uint32_t nq;
    size_t q_stride;
    uint32_t nchunks;
    const float* xn;
    const float* qn;
    uint32_t groups, KP, k;
    float eps_rel;
    const float* cand_d;
    const uint32_t* cand_i;
    const float* dropped_lb;
    uint64_t* out_keys;
    float* out_dists;
    uint8_t* unsafe; // [nq] 1 = answer this query with the SIMT kernels instead
    uint32_t* unsafe_count;
};

__device__ __forceinline__ bool closer_rr(float d, uint32_t id, float od, uint32_t oid) { return d < od || (d == od && id < oid); }

template <int DM> __global__ void __launch_bounds__(kRrThreads) exact_tc_rerank_kernel(const RrParams p) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const uint32_t q = blockIdx.x;
    const uint32_t C = p.groups * p.KP;
    uint32_t P2 = 1;
    while (P2 < C)
        P2 <<= 1;
    float* ub = reinterpret_cast<float*>(sm_raw);               // [P2] upper bounds, sorted to find the k-th
    float* sel_d = ub + P2;                                     // [kRrCap]
    uint32_t* sel_i = reinterpret_cast<uint32_t*>(sel_d + kRrCap); // [kRrCap]
    uint4* sq = reinterpret_cast<uint4*>(sel_i + kRrCap);       // [nchunks]
    __shared__ uint32_t nsel;
    __shared__ float kth_ub_s, sa2;
    __shared__ int bad;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float qn = p.qn[q];
    const float* cd = p.cand_d + (size_t)q * C;
    const uint32_t* ci = p.cand_i + (size_t)q * C;
    if (threadIdx.x == 0)
        nsel = 0, bad = 0;
    for (uint32_t i = threadIdx.x; i < P2; i += blockDim.x) {
        float u = INFINITY;
        if (i < C && ci[i] != 0xFFFFFFFFu)
            u = cd[i] + approx_eps<DM>(p.eps_rel, qn, __ldg(p.xn + ci[i]));
        ub[i] = u;
    }
    for (uint32_t c = threadIdx.x; c < p.nchunks; c += blockDim.x)
        sq[c] = __ldg(reinterpret_cast<const uint4*>(p.queries + (size_t)q * p.q_stride) + c);
    __syncthreads();
    // bitonic sort of the upper bounds (ascending)
    for (uint32_t kk = 2; kk <= P2; kk <<= 1)
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < P2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const float a = ub[i], b = ub[ixj];
                    const bool up = (i & kk) == 0;
                    if ((a > b) == up)
                        ub[i] = b, ub[ixj] = a;
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0)
        kth_ub_s = p.k <= P2 ? ub[p.k - 1] : INFINITY; // +inf when fewer than k candidates exist: everything qualifies
    __syncthreads();
    const float kth_ub = kth_ub_s;
    // anything a list dropped that could still qualify makes the query unsafe
    for (uint32_t gI = threadIdx.x; gI < p.groups; gI += blockDim.x)
        if (p.dropped_lb[(size_t)q * p.groups + gI] <= kth_ub)
            bad = 1;
    // survivors
    for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) {
        const uint32_t id = ci[i];
        if (id == 0xFFFFFFFFu)
            continue;
        const float lb = cd[i] - approx_eps<DM>(p.eps_rel, qn, __ldg(p.xn + id));
        if (lb <= kth_ub) {
            const uint32_t at = atomicAdd(&nsel, 1u);
            if (at < kRrCap)
                sel_i[at] = id;
            else
                bad = 1;
        }
    }
    __syncthreads();
    const uint32_t ns = min(nsel, kRrCap);
    if (bad || ns < min((size_t)p.k, p.n)) {
        if (threadIdx.x == 0) {
            p.unsafe[q] = 1;
            atomicAdd(p.unsafe_count, 1u);
        }
        return;
    }
    if (threadIdx.x == 0)
        p.unsafe[q] = 0;
    // exact distances: one warp per survivor, exact.cu's lane layout (lane l owns chunks l, l+32, ...)
    if (DM == DM_COS) {
        if (warp == 0) {
            float part = 0.f;
            for (uint32_t c = lane; c < p.nchunks; c += 32)
                part = norm_add(part, query_norm_chunk<DM, SK_F32>(sq[c]));
            part = warp_sum(part);
            if (lane == 0)
                sa2 = part;
        }
        __syncthreads();
    }
    for (uint32_t s = warp; s < ns; s += kRrThreads / 32) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.data + (size_t)sel_i[s] * p.data_stride);
        DistAcc<DM, SK_F32> acc;
        acc.reset();
        for (uint32_t c = lane; c < p.nchunks; c += 32)
            accum_chunk<DM, SK_F32>(acc, sq[c], __ldg(rp + c));
        const float d = finish_distance<DM, SK_F32>(acc, DM == DM_COS ? sa2 : 0.f);
        if (lane == 0)
            sel_d[s] = d;
    }
    __syncthreads();
    // rank by (distance, offset): the survivor at rank r < k is output r.  O(ns^2 / threads), ns is a few dozen.
    for (uint32_t s = threadIdx.x; s < ns; s += blockDim.x) {
        const float d = sel_d[s];
        const uint32_t id = sel_i[s];
        uint32_t rank = 0;
        for (uint32_t t = 0; t < ns; ++t)
            rank += closer_rr(sel_d[t], sel_i[t], d, id) ? 1u : 0u;
        if (rank < p.k) {
            p.out_keys[(size_t)q * p.k + rank] = id;
            p.out_dists[(size_t)q * p.k + rank] = d;
        }
    }
    for (uint32_t r = ns + threadIdx.x; r < p.k; r += blockDim.x) { // fewer rows than k
        p.out_keys[(size_t)q * p.k + r] = ~0ull;
        p.out_dists[(size_t)q * p.k + r] = INFINITY;
    }
}

} // namespace

bool exact_tc_applicable(int dist_mode, int scalar_kind, size_t n, size_t nq, size_t k, uint32_t row_bytes) {
    if (const char* e = getenv("LB200_EXACT")) {
        if (e[0] == 's') // "simt"
            return false;
        if (e[0] == 't') // "tc": force (still only where it is defined)
            return scalar_kind == SK_F32 && (dist_mode == DM_L2SQ || dist_mode == DM_COS) && k <= 192 && n < 0xFFFFFFFFull && n && nq;
    }
    // the tensor-core pass pays off once the problem is a real GEMM
    return scalar_kind == SK_F32 && (dist_mode == DM_L2SQ || dist_mode == DM_COS) && k <= 192 && n >= 32768 && n < 0xFFFFFFFFull &&
           nq >= 16 && row_bytes >= 128;
}

// Returns the device array of per-query "unsafe" flags (owned by the caller through `scratch`, freed with cudaFreeAsync by the
// caller) so that launch_exact can run the SIMT kernels for exactly those queries.
void launch_exact_tc(int dist_mode, const uint8_t* d_data, size_t n, size_t data_stride, const uint8_t* d_queries, size_t nq,
                     size_t q_stride, uint32_t row_bytes, size_t k, uint64_t* d_keys, float* d_dists, uint8_t** d_unsafe,
                     uint32_t** d_unsafe_count, cudaStream_t stream) {
    const uint32_t nchunks = row_bytes / 16, dims = row_bytes / 4;
    int sms = device_sm_count();
    const uint32_t q_tiles = (uint32_t)((nq + kTileM - 1) / kTileM);
    const uint32_t row_tiles = (uint32_t)((n + kTileN - 1) / kTileN);
    uint32_t groups = std::max<uint32_t>(1, (uint32_t)sms / q_tiles);
    groups = std::min(groups, row_tiles);
    uint32_t KP = (uint32_t)round_up(k + std::max<size_t>(32, k / 2), 32);
    KP = std::min(KP, kMaxKP);
    const float eps_rel = (float)(dims + 32) * 5.9604645e-8f; // (dims + 32) * 2^-24

    float *xn = nullptr, *qn = nullptr, *cand_d = nullptr, *dropped = nullptr;
    uint32_t* cand_i = nullptr;
    LB_CUDA(cudaMallocAsync(&xn, n * sizeof(float), stream));
    LB_CUDA(cudaMallocAsync(&qn, nq * sizeof(float), stream));
    LB_CUDA(cudaMallocAsync(&cand_d, nq * groups * KP * sizeof(float), stream));
    LB_CUDA(cudaMallocAsync(&cand_i, nq * groups * KP * sizeof(uint32_t), stream));
    LB_CUDA(cudaMallocAsync(&dropped, nq * groups * sizeof(float), stream));
    LB_CUDA(cudaMallocAsync(d_unsafe, nq, stream));
    LB_CUDA(cudaMallocAsync(d_unsafe_count, sizeof(uint32_t), stream));
    LB_CUDA(cudaMemsetAsync(*d_unsafe_count, 0, sizeof(uint32_t), stream));
    norms_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(d_data, n, data_stride, nchunks, xn);
    norms_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(d_queries, nq, q_stride, nchunks, qn);
    count_launch(2);

    TcParams p{};
    p.data = d_data, p.n = n, p.data_stride = data_stride;
    p.queries = d_queries, p.nq = (uint32_t)nq, p.q_stride = q_stride, p.nchunks = nchunks;
    p.xn = xn, p.qn = qn, p.groups = groups, p.row_tiles = row_tiles, p.KP = KP, p.eps_rel = eps_rel;
    p.cand_d = cand_d, p.cand_i = cand_i, p.dropped_lb = dropped;
    RrParams r{};
    r.data = d_data, r.n = n, r.data_stride = data_stride, r.queries = d_queries, r.nq = (uint32_t)nq, r.q_stride = q_stride;
    r.nchunks = nchunks, r.xn = xn, r.qn = qn, r.groups = groups, r.KP = KP, r.k = (uint32_t)k, r.eps_rel = eps_rel;
    r.cand_d = cand_d, r.cand_i = cand_i, r.dropped_lb = dropped, r.out_keys = d_keys, r.out_dists = d_dists;
    r.unsafe = *d_unsafe, r.unsafe_count = *d_unsafe_count;
    uint32_t P2 = 1;
    while (P2 < groups * KP)
        P2 <<= 1;
    const size_t rr_smem = (size_t)P2 * 4 + (size_t)kRrCap * 8 + row_bytes;
    auto run = [&](auto dm) {
        constexpr int DM = decltype(dm)::value;
        auto fk = exact_tc_filter_kernel<DM>;
        LB_CUDA(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        fk<<<q_tiles * groups, kTcThreads, kTcSmem, stream>>>(p);
        LB_CUDA(cudaGetLastError());
        auto rk = exact_tc_rerank_kernel<DM>;
        LB_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));
        rk<<<(unsigned)nq, kRrThreads, rr_smem, stream>>>(r);
        LB_CUDA(cudaGetLastError());
        count_launch(2);
    };
    if (dist_mode == DM_COS)
        run(std::integral_constant<int, DM_COS>{});
    else
        run(std::integral_constant<int, DM_L2SQ>{});
    LB_CUDA(cudaFreeAsync(xn, stream));
    LB_CUDA(cudaFreeAsync(qn, stream));
    LB_CUDA(cudaFreeAsync(cand_d, stream));
    LB_CUDA(cudaFreeAsync(cand_i, stream));
    LB_CUDA(cudaFreeAsync(dropped, stream));
}

} // namespace lb200
