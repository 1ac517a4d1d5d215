// lantern_b200 -- host-side index object behind the C ABI (include/lantern_b200.h).
//
// HBM layout of one index (one shard = one GPU):
//   vectors   [capacity][row_bytes]   stored scalar kind (f32/f16/i8/b1), row_bytes = round_up(bytes,16),
//                                     zero padded; base 256-B aligned -> every row is a whole number of
//                                     16-B chunks and (for d=768 f32) of 128-B lines.   PQ: codes[capacity][nsub_pad]
//   adj0      [capacity][M0] u32      level-0 adjacency, stored order, padded with 0xFFFFFFFF
//   upper_ref [capacity] u32          0xFFFFFFFF for level-0-only nodes, else offset/M into upper_adj of the node's
//                                     level-1 list; level l list at (ref + l-1)*M
//   upper_adj [sum(levels)*M] u32     padded with 0xFFFFFFFF
//   keys      [capacity] u64, levels [capacity] i16
//   scratch   per resident CTA: visited bitmap (capacity bits) + touched-word list
// The reference keeps the same information in per-node byte tapes with 6-byte slots
// (U/include/usearch/index.hpp:1799-1863); only the (de)serialiser here speaks that format.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace lb200 {

struct IndexConfig {
    int metric_kind = MK_L2SQ;
    int scalar_kind = SK_F32;
    size_t dims = 0;
    size_t M = 16, M0 = 32, efc = 128, ef = 64;
    bool pq = false;
    size_t num_centroids = 0, num_subvectors = 0;
};

struct SearchStats {
    uint64_t queries = 0, computed_distances = 0, base_pops = 0, upper_hops = 0, algorithmic_bytes = 0, limbo_overflows = 0;
    double kernel_ms = 0; // device time of the search kernel alone (CUDA events on the launching stream)
};

// Device-side view handed to kernels (plain pointers, trivially copyable).
struct GraphView {
    const uint8_t* vectors;
    const uint32_t* adj0;
    const uint32_t* upper_ref;
    const uint32_t* upper_adj;
    const uint64_t* keys;
    uint32_t n;
    uint32_t row_bytes; // bytes per stored row (multiple of 16)
    uint32_t M, M0;
    uint32_t entry;
    int32_t max_level;
    // PQ (row_bytes then is the padded code width)
    const float* codebook; // [num_centroids][dims]
    const float* pq_pair;  // [nsub][ncent][ncent] centroid-pair table (l2sq: |ca-cb|^2 ; cos: ca.cb)
    const float* pq_norm;  // [nsub][ncent] |centroid slice|^2 (cos)
    const float* pq_query_tables; // search only: per-query look-up tables [nq][pq_table_floats] precomputed for the batch, or NULL
    uint32_t flags;        // tuning: 1 = L2-prefetch the adjacency of every measured node, 2 = of accepted nodes only
    uint32_t dims, num_centroids, num_subvectors;
    uint32_t pq_lut_width; // centroids the per-value table covers (= num_centroids, or 128 when every stored code is < 128)
};

struct SearchScratch {
    uint32_t* visited;     // [ctas][words_per_cta]
    uint32_t* touched;     // [ctas][touched_cap]
    unsigned long long* counters; // [0]=next query, [1]=dist evals, [2]=base pops, [3]=upper hops, [4]=limbo overflows (8 slots)
    size_t words_per_cta;
    uint32_t touched_cap;
    uint32_t ctas;
};

class Index {
  public:
    explicit Index(const IndexConfig& cfg, const float* codebook);
    ~Index();

    const IndexConfig& config() const { return cfg_; }
    size_t size() const { return n_ + pending_n_ + staged_keys_.size(); }
    size_t capacity() const { return capacity_; }
    size_t row_bytes() const { return row_bytes_; }

    void reserve(size_t capacity);
    // staging (host or device source) of vectors in the *input* kind (f32 or b1)
    void add_host(const uint64_t* keys, const void* vectors, size_t n, size_t stride, int kind);
    void add_device(const uint64_t* host_keys, const void* d_vectors, size_t n, size_t stride, int kind);
    void add_one_host(uint64_t key, const void* vector, int kind); // usearch_add: staged on the host, flushed in blocks
    void flush_staged();                                           // caller holds no lock
    void build(); // insert all pending vectors (build.cu)

    // search: queries in device memory, input kind f32 or b1
    void search_device(const void* d_queries, size_t nq, size_t stride, int kind, size_t k, size_t ef, uint64_t* d_keys,
                       float* d_dists, uint32_t* d_counts, cudaStream_t stream);
    void search_host(const void* queries, size_t nq, size_t stride, int kind, size_t k, size_t ef, uint64_t* keys,
                     float* dists, size_t* counts);
    SearchStats last_stats();

    // usearch/lantern file format (format.cc)
    size_t serialized_length();
    size_t serialized_length_locked(); // caller holds mu_
    size_t save_buffer(void* buffer, size_t length);
    void load_buffer(const void* buffer, size_t length);
    void write_header(void* headerp);      // 136 bytes (usearch_update_header)
    void fill_header(uint8_t* p) const;    // caller holds mu_
    size_t count_key(uint64_t key);        // usearch_count: vectors stored under `key`

    GraphView view() const;

    // --- used by build.cu / format.cc ---
    void ensure_capacity(size_t cap);
    void alloc_upper(size_t total_lists);

    IndexConfig cfg_;
    int dist_mode_ = 0;
    size_t row_bytes_ = 0;    // stored row (or padded PQ code) bytes
    size_t vec_bytes_ = 0;    // unpadded bytes of a vector in the metric scalar kind (file format)
    size_t stored_bytes_ = 0; // unpadded bytes kept per node in the file (vec_bytes_ or num_subvectors)
    size_t n_ = 0, capacity_ = 0;
    int32_t max_level_ = -1;
    uint32_t entry_ = 0;
    uint32_t level_rng_ = 1u;   // minstd_rand0 state, default seed (std::default_random_engine, index.hpp:2082)
    size_t build_batch_ = 0;    // max nodes inserted per batch; 0 = one per resident CTA; 1 = the reference's sequential order
    size_t build_ratio_ = 64;   // a batch never exceeds (visible nodes) / build_ratio_
    uint32_t touched_cap_ = 16384; // per-CTA log of bitmap words to un-visit; beyond it the whole bitmap is cleared
    size_t search_expand_ = 1;  // 1 = exact-order search; 2..4 = relaxed order (see walk.cuh)
    size_t search_kernel_ = 0;  // 0 = pick by row width, 1 = one CTA per query (search.cu), 2 = one warp per query (group.cu)
    uint8_t* d_warp_aux_ = nullptr; // warp kernel: counters, flags, dummy counts
    size_t warp_aux_bytes_ = 0;
    double last_build_ms_ = 0;
    uint64_t last_build_dist_ = 0;
    size_t last_build_n_ = 0;

    // device arrays
    uint8_t* d_vectors_ = nullptr;
    uint32_t* d_adj0_ = nullptr;
    uint32_t* d_upper_ref_ = nullptr;
    uint32_t* d_upper_adj_ = nullptr;
    size_t upper_lists_ = 0, upper_lists_cap_ = 0; // number of M-wide lists in d_upper_adj_
    uint64_t* d_keys_ = nullptr;
    float* d_codebook_ = nullptr;
    float* d_pq_pair_ = nullptr;
    float* d_pq_norm_ = nullptr;
    float* d_pq_tables_ = nullptr; // per-query look-up tables of the current search batch
    size_t pq_tables_bytes_ = 0;
    uint32_t pq_max_code_ = 0; // largest centroid id present in the stored codes
    float* d_pending_raw_ = nullptr; // pq: raw f32 rows of the pending vectors (the value side of build distances)
    size_t pending_raw_cap_ = 0;
    // host mirrors of the small per-node metadata
    std::vector<int16_t> h_levels_;
    std::vector<uint64_t> h_keys_;

    // pending (not yet inserted) vectors live directly in d_vectors_[n_ ...); only bookkeeping here
    size_t pending_n_ = 0;
    // vectors added one at a time (usearch_add, build.c:128) wait here until a block is worth a host->device copy
    std::vector<uint8_t> staged_rows_;
    std::vector<uint64_t> staged_keys_;
    int staged_kind_ = 0;
    std::mutex stage_mu_;

    // search scratch
    SearchScratch scratch_{};
    uint8_t* d_query_buf_ = nullptr;
    size_t query_buf_bytes_ = 0;
    void* d_io_buf_ = nullptr;
    size_t io_buf_bytes_ = 0;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    uint32_t last_nq_ = 0;
    std::mutex mu_;
    std::recursive_mutex host_mu_; // whole host-buffer calls (they share d_io_buf_); always taken before mu_
    // streaming state of the single-query entry point (usearch_search_ef continue_search, scan.c:240-292)
    std::vector<uint8_t> stream_query_;
    std::vector<uint64_t> stream_returned_; // keys handed out so far for stream_query_
    std::mutex stream_mu_;

    void ensure_scratch(uint32_t ctas);
    void* io_buffer(size_t bytes);
    uint8_t* query_buffer(size_t bytes);
};

// ---- kernels' host launchers -----------------------------------------------------------------------
// codec.cu
void launch_cast_rows(const void* d_in, size_t in_stride, int in_kind, void* d_out, size_t out_stride, int out_kind,
                      size_t dims, size_t n, cudaStream_t stream);
void launch_pq_encode(const float* d_codebook, size_t dims, size_t ncent, size_t nsub, const float* d_vecs,
                      size_t vec_stride_floats, size_t n, uint8_t* d_codes, size_t code_stride, bool compat128,
                      cudaStream_t stream);
void launch_pq_tables(const float* d_codebook, size_t dims, size_t ncent, size_t nsub, bool cosine, float* d_pair, float* d_norm,
                      cudaStream_t stream);
// per-query ADC tables of a batch: out[q][s * lut_width + c] (+ |q|^2 at [nsub * lut_width]), row stride = nsub*lut_width + 4 floats
void launch_pq_query_tables(const float* d_codebook, size_t dims, size_t nsub, size_t lut_width, bool cosine, const float* d_queries,
                            size_t q_stride_floats, size_t nq, float* d_tables, cudaStream_t stream);
void launch_pq_decode(const float* d_codebook, size_t dims, size_t ncent, size_t nsub, const uint8_t* d_codes,
                      size_t code_stride, size_t n, float* d_vecs, cudaStream_t stream);
// search.cu
struct SearchLaunch {
    GraphView g;
    SearchScratch s;
    const uint8_t* queries; // storage kind, stride = query_stride
    uint32_t query_stride;
    uint32_t nq, k, L;
    uint32_t expand; // candidates expanded per round (1 = the reference's exact order)
    uint64_t* out_keys;
    float* out_dists;
    uint32_t* out_counts;
};
uint32_t search_max_ctas(int dist_mode, int scalar_kind, const GraphView& g, uint32_t L, bool pq, uint32_t expand);
void launch_search(int dist_mode, int scalar_kind, bool pq, const SearchLaunch& p, cudaStream_t stream);
// group.cu: the warp-per-query kernel as a single-GPU search path
void launch_warp_search(Index& idx, const uint8_t* qbuf, size_t qrow, size_t nq, size_t k, uint32_t L, uint64_t* d_keys, float* d_dists,
                        uint32_t* d_counts, cudaStream_t stream);
// exact.cu
void launch_exact(int dist_mode, int scalar_kind, const uint8_t* d_data, size_t n, size_t data_stride,
                  const uint8_t* d_queries, size_t nq, size_t q_stride, uint32_t row_bytes, size_t k, uint64_t* d_keys,
                  float* d_dists, cudaStream_t stream);
void launch_pair_distance(int dist_mode, int scalar_kind, const uint8_t* d_a, size_t a_stride, const uint8_t* d_b,
                          size_t b_stride, size_t n, uint32_t row_bytes, float* d_out, cudaStream_t stream);
void launch_merge_shards(const uint64_t* d_keys, const float* d_dists, size_t shards, size_t nq, size_t k,
                         uint64_t* d_out_keys, float* d_out_dists, cudaStream_t stream);
// build.cu
void build_pending(Index& idx);
// kmeans.cu
int train_pq_codebook(const float* d_vectors, size_t stride_floats, size_t n, size_t dims, size_t nsub, size_t ncent, bool cosine,
                      size_t max_iter, uint64_t seed, const uint32_t* init_rows, float* d_codebook, cudaStream_t stream);

int device_sm_count();
void require_device();

} // namespace lb200
