// nb_probe.cu -- empirical FMA-pipe peaks of the device (the denominators the force kernels are
// judged against): dependent-chain-free FFMA, packed FFMA2 and DFMA loops, timed with CUDA events.
#include "nb_common.cuh"

namespace nb {

template <int kMode>  // 0: FFMA, 1: FFMA2 (f32x2), 2: DFMA
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float seed) {
    if (kMode == 0) {
        float a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = seed + k;
        const float b = seed * 0.5f, c = seed * 0.25f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], b, c);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += a[k];
        if (s == 123.456f) out[0] = s;
    } else if (kMode == 1) {
        float2 a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = make_float2(seed + k, seed - k);
        const float2 b = make_float2(seed * 0.5f, seed * 0.75f), c = make_float2(seed * 0.25f, seed * 0.125f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __ffma2_rn(a[k], b, c);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += a[k].x + a[k].y;
        if (s == 123.456f) out[0] = s;
    } else {
        double a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = (double)seed + k;
        const double b = seed * 0.5, c = seed * 0.25;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = fma(a[k], b, c);
        }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += a[k];
        if (s == 123.456) out[0] = (float)s;
    }
}

}  // namespace nb

extern "C" int nb_probe_fma_peak(int mode, double* tflops, float* scratch, nb_stream_t s) {
    NB_REQUIRE(mode >= 0 && mode <= 2 && tflops && scratch, "nb_probe_fma_peak: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    int dev = 0, sms = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int iters = 8192, grid = sms * 8, block = 256;
    cudaEvent_t e0, e1;
    NB_CUDA_OK(cudaEventCreate(&e0));
    NB_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        NB_CUDA_OK(cudaEventRecord(e0, st));
        if (mode == 0) nb::fma_probe_kernel<0><<<grid, block, 0, st>>>(scratch, iters, 1.0001f);
        else if (mode == 1) nb::fma_probe_kernel<1><<<grid, block, 0, st>>>(scratch, iters, 1.0001f);
        else nb::fma_probe_kernel<2><<<grid, block, 0, st>>>(scratch, iters, 1.0001f);
        NB_CUDA_OK(cudaEventRecord(e1, st));
        NB_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        NB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double fmas = (double)grid * block * iters * 16 * (mode == 1 ? 2 : 1);
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    return nb::check_launch("fma probe");
}

// ---------------------------------------------------------------------------------------------------------------
// nb_probe_occupy: n_ctas thread blocks that each hold smem_bytes of shared memory and spin for `milliseconds`.
// A stand-in for "somebody else's kernel is on the GPU" in the tests of the launches that need co-residency.
// ---------------------------------------------------------------------------------------------------------------
namespace nb {
__global__ void occupy_kernel(unsigned long long ns) {
    extern __shared__ char hold[];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) hold[0] = 1;
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}
}  // namespace nb

extern "C" int nb_probe_occupy(int n_ctas, size_t smem_bytes, double milliseconds, nb_stream_t s) {
    NB_REQUIRE(n_ctas > 0 && milliseconds >= 0.0 && milliseconds <= 5000.0, "nb_probe_occupy: bad argument");
    NB_CUDA_OK(cudaFuncSetAttribute(nb::occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    nb::occupy_kernel<<<n_ctas, 32, smem_bytes, (cudaStream_t)s>>>((unsigned long long)(milliseconds * 1e6));
    return nb::check_launch("occupy kernel");
}
