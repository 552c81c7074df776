// exp_f64.cu -- development experiment (not part of the library): tile shape / unroll / occupancy sweep of the
// float64 force loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/exp_f64 tools/exp_f64.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../nbody-gnn-hpc_b200/csrc/nb_common.cuh"
namespace nb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 2; } }
using namespace nb;
constexpr int kStages = 4, kTileBytes = 4096;

template <class Consume>
__device__ __forceinline__ void stream_tiles(const char* __restrict__ src, int total_bytes, char* ring, uint64_t* bars, Consume&& consume) {
    const int n_tiles = (total_bytes + kTileBytes - 1) / kTileBytes;
    if (threadIdx.x == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1); mbar_init_fence(); }
    __syncthreads();
    if (threadIdx.x == 0) for (int t = 0; t < kStages && t < n_tiles; ++t) {
        const int bytes = min(kTileBytes, total_bytes - t * kTileBytes);
        mbar_arrive_expect_tx(&bars[t], bytes); tma_load_1d(ring + t * kTileBytes, src + (size_t)t * kTileBytes, bytes, &bars[t]);
    }
    for (int t = 0; t < n_tiles; ++t) {
        const int slot = t % kStages;
        mbar_wait(&bars[slot], (t / kStages) & 1);
        const int bytes = min(kTileBytes, total_bytes - t * kTileBytes);
        consume(ring + slot * kTileBytes, bytes);
        __syncthreads();
        const int nt = t + kStages;
        if (threadIdx.x == 0 && nt < n_tiles) {
            const int nbytes = min(kTileBytes, total_bytes - nt * kTileBytes);
            mbar_arrive_expect_tx(&bars[slot], nbytes); tma_load_1d(ring + slot * kTileBytes, src + (size_t)nt * kTileBytes, nbytes, &bars[slot]);
        }
    }
}

template <int kP, int kBlock, int kUnroll, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
force(const double* __restrict__ stream, int n_pad, int n_i, int seg_len, double eps2, double* __restrict__ partial) {
    __shared__ __align__(128) char ring[kStages * kTileBytes];
    __shared__ __align__(8) uint64_t bars[kStages];
    const int seg = blockIdx.y, j0 = seg * seg_len, j1 = min(j0 + seg_len, n_pad);
    const int li0 = blockIdx.x * (kBlock * kP) + threadIdx.x;
    double xi[kP], yi[kP], zi[kP], ax[kP], ay[kP], az[kP];
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const double4 q = reinterpret_cast<const double4*>(stream)[min(li0 + k * kBlock, n_i - 1)];
        xi[k] = q.x; yi[k] = q.y; zi[k] = q.z; ax[k] = ay[k] = az[k] = 0.0;
    }
    auto consume = [&](const char* tile, int bytes) {
        const double2* __restrict__ t = reinterpret_cast<const double2*>(tile);
        const int n_j = bytes >> 5;
#pragma unroll kUnroll
        for (int j = 0; j < n_j; ++j) {
            const double2 a = t[2 * j], b = t[2 * j + 1];
#pragma unroll
            for (int k = 0; k < kP; ++k) pair_f64<false>(xi[k], yi[k], zi[k], a.x, a.y, b.x, b.y, eps2, ax[k], ay[k], az[k]);
        }
    };
    stream_tiles(reinterpret_cast<const char*>(stream) + (size_t)j0 * 32, (j1 - j0) * 32, ring, bars, consume);
    double* __restrict__ out = partial + (size_t)seg * 3 * n_i;
#pragma unroll
    for (int k = 0; k < kP; ++k) { const int li = li0 + k * kBlock; if (li < n_i) { out[li] = ax[k]; out[(size_t)n_i + li] = ay[k]; out[(size_t)2 * n_i + li] = az[k]; } }
}

static double* d_stream; static double* d_partial; static int N, NSEG, SEGLEN;
template <int kP, int kBlock, int kUnroll, int kMinB>
void bench() {
    dim3 grid((N + kP * kBlock - 1) / (kP * kBlock), NSEG);
    auto k = force<kP, kBlock, kUnroll, kMinB>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kBlock, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<<<grid, kBlock>>>(d_stream, N, N, SEGLEN, 1e-4, d_partial); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
    }
    printf("P=%d block=%d unroll=%d minb=%d regs=%3d occ=%d warps/SM=%2d ctas=%5d  %.4f ms  %.1f Gint/s  (%.1f%% of 1861)\n", kP, kBlock, kUnroll, kMinB,
           fa.numRegs, occ, occ * kBlock / 32, grid.x * grid.y, best, (double)N * (N - 1) / best / 1e6, (double)N * (N - 1) / best / 1e6 / 18.61);
    fflush(stdout);
}
int main(int argc, char** argv) {
    N = argc > 1 ? atoi(argv[1]) : 65536; SEGLEN = argc > 2 ? atoi(argv[2]) : 4096; NSEG = (N + SEGLEN - 1) / SEGLEN;
    std::vector<double> h((size_t)N * 4); srand(1);
    for (int i = 0; i < N; ++i) { for (int c = 0; c < 3; ++c) h[(size_t)i * 4 + c] = (double)rand() / RAND_MAX; h[(size_t)i * 4 + 3] = 1.0 / N; }
    cudaMalloc(&d_stream, h.size() * 8); cudaMemcpy(d_stream, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&d_partial, (size_t)NSEG * 3 * N * 8);
    printf("N=%d seg_len=%d n_seg=%d\n", N, SEGLEN, NSEG);
    bench<2, 256, 4, 1>(); bench<2, 256, 4, 2>(); bench<2, 256, 4, 3>(); bench<2, 256, 4, 4>();
    bench<2, 256, 2, 2>(); bench<2, 256, 2, 4>(); bench<2, 256, 1, 4>(); bench<2, 256, 8, 2>();
    bench<1, 256, 4, 4>(); bench<1, 256, 4, 8>(); bench<1, 256, 8, 4>(); bench<1, 128, 4, 8>(); bench<1, 128, 8, 16>();
    bench<3, 256, 2, 2>(); bench<3, 256, 4, 2>(); bench<4, 256, 2, 2>(); bench<4, 256, 2, 1>(); bench<4, 128, 2, 4>(); bench<4, 128, 4, 2>();
    bench<2, 128, 4, 4>(); bench<2, 128, 4, 8>(); bench<2, 128, 2, 8>(); bench<2, 512, 4, 1>(); bench<2, 512, 2, 2>();
    return 0;
}
