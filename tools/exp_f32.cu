// exp_f32.cu -- development experiment (not part of the library): which mix of packed (f32x2) and
// scalar FP32 instructions, i-bodies per thread and block size gives the fastest float32 force loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/exp_f32 tools/exp_f32.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../nbody-gnn-hpc_b200/csrc/nb_common.cuh"

namespace nb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 2; } }
extern "C" { int nb_padded_bodies(int n) { return (n + 31) / 32 * 32; }
int nb_segment_plan(int, int*, int*) { return 0; }
size_t nb_workspace_bytes(int, int, int) { return 0; } }
#define kStages kLibStages
#define kTileBytes kLibTileBytes
#define stream_tiles lib_stream_tiles
#include "../nbody-gnn-hpc_b200/csrc/nb_force.cu"   // the library kernels, timed beside the variants below
#undef kStages
#undef kTileBytes
#undef stream_tiles
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

template <bool P> __device__ __forceinline__ float2 add2(float2 a, float2 b) { if (P) return __fadd2_rn(a, b); return make_float2(a.x + b.x, a.y + b.y); }
template <bool P> __device__ __forceinline__ float2 mul2(float2 a, float2 b) { if (P) return __fmul2_rn(a, b); return make_float2(a.x * b.x, a.y * b.y); }
template <bool P> __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { if (P) return __ffma2_rn(a, b, c); return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }

// kPack bits: 1 = SUB group packed, 2 = R2 group, 4 = MUL group, 8 = ACC group
template <int kP, int kBlock, int kPack, int kUnroll, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
force(const float* __restrict__ stream, int n_pad, int n_i, int seg_len, float eps2, float* __restrict__ partial) {
    __shared__ __align__(128) char ring[kStages * kTileBytes];
    __shared__ __align__(8) uint64_t bars[kStages];
    const int seg = blockIdx.y, j0 = seg * seg_len, j1 = min(j0 + seg_len, n_pad);
    const int li0 = blockIdx.x * (kBlock * kP) + threadIdx.x;
    float xi[kP], yi[kP], zi[kP]; float2 ax[kP], ay[kP], az[kP];
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int gi = min(li0 + k * kBlock, n_i - 1);
        xi[k] = stream[(size_t)(gi >> 1) * 8 + (gi & 1)]; yi[k] = stream[(size_t)(gi >> 1) * 8 + 2 + (gi & 1)]; zi[k] = stream[(size_t)(gi >> 1) * 8 + 4 + (gi & 1)];
        ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
    }
    const float2 e2 = make_float2(eps2, eps2);
    auto consume = [&](const char* tile, int bytes) {
        const float4* __restrict__ t = reinterpret_cast<const float4*>(tile);
        const int n_pairs = bytes >> 5;
#pragma unroll kUnroll
        for (int jp = 0; jp < n_pairs; ++jp) {
            const float4 A = t[2 * jp], B = t[2 * jp + 1];
            const float2 xj = make_float2(A.x, A.y), yj = make_float2(A.z, A.w), zj = make_float2(B.x, B.y), gj = make_float2(B.z, B.w);
#pragma unroll
            for (int k = 0; k < kP; ++k) {
                const float2 dx = add2<(kPack & 1) != 0>(xj, make_float2(-xi[k], -xi[k]));
                const float2 dy = add2<(kPack & 1) != 0>(yj, make_float2(-yi[k], -yi[k]));
                const float2 dz = add2<(kPack & 1) != 0>(zj, make_float2(-zi[k], -zi[k]));
                float2 r2 = fma2<(kPack & 2) != 0>(dx, dx, e2);
                r2 = fma2<(kPack & 2) != 0>(dy, dy, r2);
                r2 = fma2<(kPack & 2) != 0>(dz, dz, r2);
                float2 inv; inv.x = rsqrt_approx(r2.x); inv.y = rsqrt_approx(r2.y);
                const float2 inv2 = mul2<(kPack & 4) != 0>(inv, inv);
                float2 f = mul2<(kPack & 4) != 0>(gj, inv);
                f = mul2<(kPack & 4) != 0>(f, inv2);
                ax[k] = fma2<(kPack & 8) != 0>(f, dx, ax[k]);
                ay[k] = fma2<(kPack & 8) != 0>(f, dy, ay[k]);
                az[k] = fma2<(kPack & 8) != 0>(f, dz, az[k]);
            }
        }
    };
    stream_tiles(reinterpret_cast<const char*>(stream) + (size_t)j0 * 16, (j1 - j0) * 16, ring, bars, consume);
    float* __restrict__ out = partial + (size_t)seg * 3 * n_i;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int li = li0 + k * kBlock;
        if (li < n_i) { out[li] = ax[k].x + ax[k].y; out[(size_t)n_i + li] = ay[k].x + ay[k].y; out[(size_t)2 * n_i + li] = az[k].x + az[k].y; }
    }
}

static float* d_stream; static float* d_partial; static int N, NSEG, SEGLEN;

template <int kP, int kBlock, int kPack, int kUnroll, int kMinB>
void bench(const char* tag) {
    dim3 grid((N + kP * kBlock - 1) / (kP * kBlock), NSEG);
    auto k = force<kP, kBlock, kPack, kUnroll, kMinB>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kBlock, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        k<<<grid, kBlock>>>(d_stream, N, N, SEGLEN, 1e-4f, d_partial);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-28s P=%d block=%d pack=%2d unroll=%d regs=%3d occ=%d ctas=%5d  %.4f ms  %.1f Gint/s  (%.1f%% of 3722)%s\n", tag, kP, kBlock, kPack,
           kUnroll, fa.numRegs, occ, grid.x * grid.y, best, (double)N * (N - 1) / best / 1e6, (double)N * (N - 1) / best / 1e6 / 37.22,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

template <int kP, int kBlock>
void bench_lib(const char* tag) {
    dim3 grid((N + kP * kBlock - 1) / (kP * kBlock), NSEG);
    auto k = nb::force_f32_kernel<kP, kBlock, false>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        k<<<grid, kBlock>>>(d_stream, N, 0, N, SEGLEN, 1e-4f, d_partial, nb::PeerWait{nullptr, 0, 0}, nb::Epilogue<float>{});
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
    }
    printf("%-28s P=%d block=%d regs=%3d ctas=%5d  %.4f ms  %.1f Gint/s  (%.1f%% of 3722)\n", tag, kP, kBlock, fa.numRegs,
           grid.x * grid.y, best, (double)N * (N - 1) / best / 1e6, (double)N * (N - 1) / best / 1e6 / 37.22);
    fflush(stdout);
}

int main(int argc, char** argv) {
    N = argc > 1 ? atoi(argv[1]) : 65536;
    SEGLEN = argc > 2 ? atoi(argv[2]) : 1024;
    NSEG = (N + SEGLEN - 1) / SEGLEN;
    std::vector<float> h((size_t)N * 4);
    srand(1);
    for (int p = 0; p < N / 2; ++p) for (int c = 0; c < 8; ++c) h[(size_t)p * 8 + c] = c < 6 ? (float)rand() / RAND_MAX : 1.0f / N;
    cudaMalloc(&d_stream, h.size() * 4); cudaMemcpy(d_stream, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_partial, (size_t)NSEG * 3 * N * 4);
    printf("N=%d seg_len=%d n_seg=%d\n", N, SEGLEN, NSEG);
    bench_lib<4, 256>("LIBRARY kernel"); bench_lib<2, 128>("LIBRARY kernel");
    bench<4, 256, 15, 2, 1>("u2"); bench<4, 256, 15, 4, 1>("u4");
    if (argc > 3) return 0;
#define PACKS(P, B, U, M) \
    bench<P, B, 15, U, M>("all packed"); bench<P, B, 0, U, M>("all scalar"); bench<P, B, 1, U, M>("SUB"); bench<P, B, 2, U, M>("R2"); \
    bench<P, B, 4, U, M>("MUL"); bench<P, B, 8, U, M>("ACC"); bench<P, B, 3, U, M>("SUB+R2"); bench<P, B, 5, U, M>("SUB+MUL"); \
    bench<P, B, 9, U, M>("SUB+ACC"); bench<P, B, 6, U, M>("R2+MUL"); bench<P, B, 10, U, M>("R2+ACC"); bench<P, B, 12, U, M>("MUL+ACC"); \
    bench<P, B, 7, U, M>("SUB+R2+MUL"); bench<P, B, 11, U, M>("SUB+R2+ACC"); bench<P, B, 13, U, M>("SUB+MUL+ACC"); bench<P, B, 14, U, M>("R2+MUL+ACC");
    PACKS(4, 256, 4, 1)
    PACKS(4, 128, 4, 1)
    PACKS(2, 256, 4, 1)
    PACKS(8, 128, 2, 1)
    bench<4, 256, 15, 2, 1>("u2"); bench<4, 256, 15, 8, 1>("u8"); bench<4, 256, 5, 2, 1>("u2"); bench<4, 256, 5, 8, 1>("u8");
    bench<4, 256, 15, 4, 2>("minb2"); bench<4, 128, 15, 4, 4>("minb4"); bench<2, 128, 15, 4, 8>("minb8"); bench<2, 256, 15, 4, 4>("minb4");
    return 0;
}
