// lantern_b200 -- PQ codebook training (k-means per subvector) on the GPU.  SURVEY.md 8f-4.
//
// Restates lantern_hnsw/src/hnsw/product_quantization.c:
//   initialize_clusters :51-72    k distinct random dataset rows per subvector
//   assign_to_clusters  :82-125   nearest centre by usearch_distance (l2sq or cos), strict '<' (lowest id wins)
//   update_centers      :152-166  mean of the assigned points; an empty cluster keeps its centre
//   should_stop_iterations :173-194  stop when mean_c distance(old_c, new_c) <= 0.1
//   k_means             :207-256  at most `iter` rounds;  product_quantization :275-293 one k-means per subvector
// Output tape = what load_pq_codebook (pqtable.c:194-333) hands to usearch_init: float[centroid][dim], subvector s of
// centroid c at c*dim + s*subdim.  Sums are accumulated in fp64 (atomics) so the result does not depend on scheduling.
#include <cuda_runtime.h>
#include <math.h>

#include <vector>

#include "distance.cuh"
#include "engine.h"

namespace lb200 {

namespace {

__device__ __forceinline__ float sub_distance(const float* a, const float* b, uint32_t sd, int cosine) {
    if (!cosine) {
        float acc = 0.f;
        for (uint32_t i = 0; i < sd; ++i) {
            const float d = a[i] - b[i];
            acc += d * d;
        }
        return acc;
    }
    float ab = 0.f, a2 = 0.f, b2 = 0.f;
    for (uint32_t i = 0; i < sd; ++i)
        ab += a[i] * b[i], a2 += a[i] * a[i], b2 += b[i] * b[i];
    return cos_from_parts(ab, a2, b2);
}

// one thread per (vector, subvector): nearest centre, then fp64 accumulation of the member sums
__global__ void kmeans_assign_kernel(const float* __restrict__ data, size_t stride_floats, size_t n, uint32_t dims, uint32_t nsub,
                                     uint32_t ncent, int cosine, const float* __restrict__ centers /*[ncent][dims]*/,
                                     const int* __restrict__ active, double* __restrict__ sums /*[nsub][ncent][sd]*/,
                                     unsigned int* __restrict__ counts /*[nsub][ncent]*/) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nsub)
        return;
    const uint32_t s = (uint32_t)(t % nsub);
    const size_t i = t / nsub;
    if (!active[s])
        return;
    const uint32_t sd = dims / nsub;
    const float* x = data + i * stride_floats + (size_t)s * sd;
    float best = 3.402823466e+38f;
    uint32_t bc = 0;
    for (uint32_t c = 0; c < ncent; ++c) {
        const float d = sub_distance(x, centers + (size_t)c * dims + (size_t)s * sd, sd, cosine);
        if (d < best)
            best = d, bc = c;
    }
    double* dst = sums + ((size_t)s * ncent + bc) * sd;
    for (uint32_t j = 0; j < sd; ++j)
        atomicAdd(dst + j, (double)x[j]);
    atomicAdd(counts + (size_t)s * ncent + bc, 1u);
}

// one thread per (subvector, centre): new centre, shift accumulated per subvector
__global__ void kmeans_update_kernel(uint32_t dims, uint32_t nsub, uint32_t ncent, int cosine, float* __restrict__ centers,
                                     const int* __restrict__ active, const double* __restrict__ sums,
                                     const unsigned int* __restrict__ counts, double* __restrict__ shift /*[nsub]*/) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsub * ncent)
        return;
    const uint32_t s = t / ncent, c = t % ncent;
    if (!active[s])
        return;
    const uint32_t sd = dims / nsub;
    float* cen = centers + (size_t)c * dims + (size_t)s * sd;
    const unsigned int cnt = counts[(size_t)s * ncent + c];
    if (!cnt)
        return; // keeps its centre, contributes distance 0 (product_quantization.c:160)
    float old[64], neu[64];
    float dist;
    if (sd <= 64) {
        for (uint32_t j = 0; j < sd; ++j) {
            old[j] = cen[j];
            neu[j] = (float)(sums[((size_t)s * ncent + c) * sd + j] / (double)cnt);
            cen[j] = neu[j];
        }
        dist = sub_distance(old, neu, sd, cosine);
    } else { // wide subvectors: two passes over global memory
        float ab = 0.f, a2 = 0.f, b2 = 0.f, l2 = 0.f;
        for (uint32_t j = 0; j < sd; ++j) {
            const float o = cen[j];
            const float v = (float)(sums[((size_t)s * ncent + c) * sd + j] / (double)cnt);
            cen[j] = v;
            ab += o * v, a2 += o * o, b2 += v * v, l2 += (o - v) * (o - v);
        }
        dist = cosine ? cos_from_parts(ab, a2, b2) : l2;
    }
    atomicAdd(shift + s, (double)dist);
}

__global__ void kmeans_init_kernel(const float* __restrict__ data, size_t stride_floats, uint32_t dims, uint32_t nsub, uint32_t ncent,
                                   const uint32_t* __restrict__ init_rows /*[nsub][ncent]*/, float* __restrict__ centers) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)ncent * dims)
        return;
    const uint32_t c = (uint32_t)(t / dims), j = (uint32_t)(t % dims);
    const uint32_t s = j / (dims / nsub);
    centers[t] = data[(size_t)init_rows[(size_t)s * ncent + c] * stride_floats + j];
}

} // namespace

// d_vectors: n rows of `dims` floats, `stride_floats` apart.  init_rows (host, [nsub][ncent]) may be null: then distinct
// rows are drawn per subvector from `seed`.  Returns the number of Lloyd rounds of the slowest subvector.
int train_pq_codebook(const float* d_vectors, size_t stride_floats, size_t n, size_t dims, size_t nsub, size_t ncent, bool cosine,
                      size_t max_iter, uint64_t seed, const uint32_t* init_rows, float* d_codebook, cudaStream_t stream) {
    if (!n || !nsub || !ncent || dims % nsub || ncent > n)
        throw CudaError("pq training: need dims % num_subvectors == 0 and at least num_centroids vectors");
    std::vector<uint32_t> rows(nsub * ncent);
    if (init_rows)
        rows.assign(init_rows, init_rows + nsub * ncent);
    else { // distinct rows per subvector (initialize_clusters), splitmix64 stream
        uint64_t st = seed ? seed : 0x9E3779B97F4A7C15ull;
        auto next = [&]() {
            uint64_t z = (st += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        };
        for (size_t s = 0; s < nsub; ++s)
            for (size_t c = 0; c < ncent; ++c) {
                for (;;) {
                    uint32_t r = (uint32_t)(next() % n);
                    bool used = false;
                    for (size_t k = 0; k < c; ++k)
                        used |= rows[s * ncent + k] == r;
                    if (!used) {
                        rows[s * ncent + c] = r;
                        break;
                    }
                }
            }
    }
    const size_t sd = dims / nsub;
    uint32_t* d_rows = nullptr;
    int* d_active = nullptr;
    double *d_sums = nullptr, *d_shift = nullptr;
    unsigned int* d_counts = nullptr;
    LB_CUDA(cudaMalloc(&d_rows, rows.size() * 4));
    LB_CUDA(cudaMalloc(&d_active, nsub * sizeof(int)));
    LB_CUDA(cudaMalloc(&d_sums, nsub * ncent * sd * sizeof(double)));
    LB_CUDA(cudaMalloc(&d_counts, nsub * ncent * sizeof(unsigned int)));
    LB_CUDA(cudaMalloc(&d_shift, nsub * sizeof(double)));
    LB_CUDA(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, stream));
    kmeans_init_kernel<<<(unsigned)((ncent * dims + 255) / 256), 256, 0, stream>>>(d_vectors, stride_floats, (uint32_t)dims, (uint32_t)nsub,
                                                                                  (uint32_t)ncent, d_rows, d_codebook);
    count_launch();
    std::vector<int> active(nsub, 1);
    std::vector<double> shift(nsub);
    int rounds = 0;
    for (size_t it = 0; it < max_iter; ++it) {
        bool any = false;
        for (int a : active)
            any |= a != 0;
        if (!any)
            break;
        ++rounds;
        LB_CUDA(cudaMemcpyAsync(d_active, active.data(), nsub * sizeof(int), cudaMemcpyHostToDevice, stream));
        LB_CUDA(cudaMemsetAsync(d_sums, 0, nsub * ncent * sd * sizeof(double), stream));
        LB_CUDA(cudaMemsetAsync(d_counts, 0, nsub * ncent * sizeof(unsigned int), stream));
        LB_CUDA(cudaMemsetAsync(d_shift, 0, nsub * sizeof(double), stream));
        const size_t total = n * nsub;
        kmeans_assign_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(d_vectors, stride_floats, n, (uint32_t)dims, (uint32_t)nsub,
                                                                                 (uint32_t)ncent, cosine ? 1 : 0, d_codebook, d_active, d_sums,
                                                                                 d_counts);
        kmeans_update_kernel<<<(unsigned)((nsub * ncent + 127) / 128), 128, 0, stream>>>((uint32_t)dims, (uint32_t)nsub, (uint32_t)ncent,
                                                                                        cosine ? 1 : 0, d_codebook, d_active, d_sums, d_counts,
                                                                                        d_shift);
        LB_CUDA(cudaGetLastError());
        count_launch(2);
        LB_CUDA(cudaMemcpyAsync(shift.data(), d_shift, nsub * sizeof(double), cudaMemcpyDeviceToHost, stream));
        LB_CUDA(cudaStreamSynchronize(stream));
        for (size_t s = 0; s < nsub; ++s)
            if (active[s] && (float)(shift[s] / (double)ncent) <= 0.1f) // should_stop_iterations
                active[s] = 0;
    }
    cudaFree(d_rows), cudaFree(d_active), cudaFree(d_sums), cudaFree(d_counts), cudaFree(d_shift);
    return rounds;
}

} // namespace lb200
