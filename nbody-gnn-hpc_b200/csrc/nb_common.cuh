// nb_common.cuh -- shared device/host helpers for the sm_100a direct-sum gravity kernels.
//
// Everything here is written for Blackwell B200 (sm_100a) only: packed f32x2
// arithmetic (FFMA2/FADD2/FMUL2), 1-D bulk TMA copies completed on mbarriers,
// MUFU.RSQ / MUFU.RSQ64H seeds.  There is no other code path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nbody_b200.h"

namespace nb {

constexpr double kG = NB_G;  // reference: src/hpc/nbody.py:18

// Bodies are streamed through shared memory in chunks of this many bodies.  Every
// j-segment, tile and padded system length is a multiple of it, so the unrolled
// inner loops never need a remainder and every bulk copy is a multiple of 16 B.
constexpr int kChunkBodies = NB_CHUNK_BODIES;

// ---------------------------------------------------------------------------------------------
// error reporting (C ABI: return codes + a thread-local message, never exceptions)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define NB_CUDA_OK(expr)                                    \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) return nb::cuda_fail(_e, #expr); \
    } while (0)

#define NB_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            nb::set_error(__VA_ARGS__);       \
            return NB_ERR_INVALID;            \
        }                                     \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return NB_OK;
}

// Workspace header of nb_accel_* / nb_step_*: one arrival counter per i-tile of the WHOLE system (smallest tile:
// 128 bodies), the finished-tiles counter and the error word of the sharded mode.  Sized by n alone.
static inline size_t ws_header_bytes(int n) {
    return (((size_t)(n > 0 ? n : 1) / 128 + 3) * sizeof(int) + 255) / 256 * 256;
}

// Behind the partials of a WHOLE-system workspace (n_i == n): the persistent multi-step kernel's words -- the grid
// barrier and one arrival counter per group of 32 bodies (nb_persist.cu).
static inline size_t ws_persist_bytes(int n) {
    return (((size_t)(n > 0 ? n : 1) / 32 + 8) * sizeof(int) + 255) / 256 * 256;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Bounds / protocol checks of our own (compute-sanitizer is closed on the pool): compiled in by -DNB_DEBUG_CHECKS
// (build.py --out lib/variants/libnb_debug.so -D NB_DEBUG_CHECKS), the GPU tests then run against that library
// (NBODY_B200_LIB).  A failed check prints where and traps, which the host sees as a launch failure.
#ifdef NB_DEBUG_CHECKS
#define NB_CHECK(cond)                                                                             \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            printf("NB_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                             \
            __trap();                                                                              \
        }                                                                                          \
    } while (0)
#else
#define NB_CHECK(cond) ((void)0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// --- mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) -------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// --- fp32 ----------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.RSQ, no denormal fix-up code
    return y;
}

// --- fp64 ----------------------------------------------------------------------------------------
// MUFU.RSQ64H seed (about 2^-22 relative).  Callers refine it; see pair_f64().
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// One softened pair interaction in fp64, accumulating into (ax, ay, az).
//   reference: src/hpc/nbody.py:47-60   r2 = dx^2+dy^2+dz^2+eps^2; factor = G*m_j / (sqrt(r2)*r2)
// gm_j = G*m_j is folded on the host side of the kernel.  factor = gm * r2^(-3/2) is formed as
//   y0 = seed(r2),  e = 1 - r2*y0^2  (one FMA on the rounded square),  r2^(-3/2) = y0^3 * (1 + 3/2 e + 15/8 e^2 + O(e^3)),
// i.e. one third-order correction of the cubed seed: truncation 35/16 e^3 < 2^-64 for |e| < 2^-21.
// 16 FP64-pipe operations + 1 MUFU per interaction.
//
// Build option -DNB_F64_PAIR_FIRST_ORDER (tools/probe_f64_variants.py measures it): first-order correction with the
// mass folded into the polynomial -- r2^(-3/2) G m ~ y0^3 (c0 + c1 p), p = r2 y0^2, c0 = 5/2 G m, c1 = -3/2 G m; the
// stream's mass slot then holds c0 and c1 = -0.6 c0 -- 14 operations (K3, c1 kept in shared memory) or 14.5 (K1/K2,
// c1 formed once per j).  Measured: 4.88 ms against 5.27 ms for the 300x200x400 ensemble, 3.81 against 4.15 ms per
// N = 65,536 evaluation.  Truncation 15/8 e^2 < 5e-13 relative, same sign for every pair: inside the 1e-10 bar, but on
// the reference's chaotic default initial conditions (e-folding ~11 steps) that is amplified to 1e-8 .. 1e-3 in the
// positions within 50 steps -- three orders of magnitude above the reference's own reordering noise, with 43 % instead
// of 97 % of the 300 data-generation systems still inside 1e-8 of the CPU restatement at step 50 (measured).  So it is an
// option for well-conditioned systems, not the default.  (A second-order polynomial with three folded constants per
// body -- 15 operations, 48-byte stream records -- was built and measured too: no faster than the 16-operation form
// below on either kernel, the extra shared-memory load per j costs what the saved operation gains.)
#ifdef NB_F64_PAIR_FIRST_ORDER
constexpr double kMassSlotF64 = 2.5;   // stream slot 3 of the f64 layout = kMassSlotF64 * G * m  (c0)
constexpr double kC1OverC0 = -0.6;     // c1 = -3/2 G m = -0.6 c0
#else
constexpr double kMassSlotF64 = 1.0;
constexpr double kC1OverC0 = 0.0;
#endif
template <bool kZeroEps>
__device__ __forceinline__ void pair_f64(double xi, double yi, double zi, double xj, double yj, double zj, double gmj,
                                         double c1j, double eps2, double& ax, double& ay, double& az) {
    const double dx = xj - xi;
    const double dy = yj - yi;
    const double dz = zj - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double y0 = rsqrt_seed(r2);
    if (kZeroEps) y0 = (r2 > 0.0) ? y0 : 0.0;  // eps == 0: the i == j term (and exact overlaps) contribute 0
#ifdef NB_F64_PAIR_FIRST_ORDER
    {
        const double s = y0 * y0;
        const double pp = r2 * s;
        const double c = fma(c1j, pp, gmj);  // gmj is c0 here
        const double t = y0 * s;
        const double f1 = t * c;
        ax = fma(f1, dx, ax);
        ay = fma(f1, dy, ay);
        az = fma(f1, dz, az);
        return;
    }
#endif
    (void)c1j;
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);  // y2 carries one rounding (2^-53): 1.5 * 2^-53 relative in f
    const double g = gmj * y0;
    const double w = y2 * g;
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    const double f = fma(w, q, w);
    ax = fma(f, dx, ax);
    ay = fma(f, dy, ay);
    az = fma(f, dz, az);
}

// NumPy-order leapfrog arithmetic: every product and sum rounded once, never contracted,
// as in  v += 0.5*dt*a ; x += dt*v  (src/hpc/nbody.py:205,208,214).
__device__ __forceinline__ double mul_add_unfused(double a, double b, double c) {
    return __dadd_rn(c, __dmul_rn(a, b));
}
__device__ __forceinline__ float mul_add_unfused(float a, float b, float c) { return __fadd_rn(c, __fmul_rn(a, b)); }

#endif  // __CUDACC__

}  // namespace nb
