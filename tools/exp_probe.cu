// exp_probe.cu -- development probe: FMA-pipe throughput as a function of how many distinct register operands an
// instruction reads (register-file bandwidth), scalar FFMA vs packed FFMA2.  Explains the ceiling of the float32
// force loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_probe tools/exp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int kMode>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
    float2 a[12], b[12], c[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { a[k] = make_float2(seed + k, seed - k); b[k] = make_float2(1.0f + 1e-7f * k, 1.0f - 1e-7f * k); c[k] = make_float2(1e-9f * k, -1e-9f * k); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            if (kMode == 0) a[k] = __ffma2_rn(a[k], b[0], c[0]);                 // 1 varying pair
            if (kMode == 1) a[k] = __ffma2_rn(a[k], b[k], c[0]);                 // 2 varying pairs
            if (kMode == 2) a[k] = __ffma2_rn(b[k], c[k], a[k]);                 // 3 varying pairs
            if (kMode == 3) { a[k].x = fmaf(a[k].x, b[0].x, c[0].x); a[k].y = fmaf(a[k].y, b[0].x, c[0].x); }   // scalar, 1 varying
            if (kMode == 4) { a[k].x = fmaf(a[k].x, b[k].x, c[0].x); a[k].y = fmaf(a[k].y, b[k].y, c[0].x); }   // scalar, 2 varying
            if (kMode == 5) { a[k].x = fmaf(b[k].x, c[k].x, a[k].x); a[k].y = fmaf(b[k].y, c[k].y, a[k].y); }   // scalar, 3 varying
            if (kMode == 6) a[k] = __fadd2_rn(a[k], make_float2(b[k].x, b[k].x));      // FADD2 pair + scalar broadcast
            if (kMode == 7) a[k] = __fmul2_rn(a[k], b[k]);                             // FMUL2 2 pairs
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += a[k].x + a[k].y + b[k].x + c[k].y;
    if (s == 123.456f) out[0] = s;
}
template <int kMode> void run(const char* what, int sms) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 8192, grid = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); probe<kMode><<<grid, 256>>>(d, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    const double lane_ops = (double)grid * 256 * iters * 12 * 2;
    const double tf = 2.0 * lane_ops / (best * 1e-3) / 1e12;
    printf("%-44s %7.2f TFLOP/s  (%.1f%% of 74.45)\n", what, tf, 100 * tf / 74.45);
    cudaFree(d);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("FFMA2, 1 varying register pair", sms); run<1>("FFMA2, 2 varying register pairs", sms); run<2>("FFMA2, 3 varying register pairs", sms);
    run<3>("FFMA,  1 varying register", sms); run<4>("FFMA,  2 varying registers", sms); run<5>("FFMA,  3 varying registers", sms);
    run<6>("FADD2, pair + broadcast scalar", sms); run<7>("FMUL2, 2 varying pairs", sms);
    return 0;
}
