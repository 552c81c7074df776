// nb_tiles.cuh -- the pieces the per-step force kernels (nb_force.cu) and the persistent multi-step kernel
// (nb_persist.cu) share, so that both produce the same bits: access to the position stream, and the inner loops that
// apply one shared-memory tile of j bodies to the kP bodies a thread holds in registers.
#pragma once

#include <type_traits>

#include "nb_common.cuh"

namespace nb {

// ------------------------------------------------------------------------------------------------
// stream layout (include/nbody_b200.h): f64 n_pad x {x, y, z, G m}; f32 pairs {x0,x1,y0,y1 | z0,z1,Gm0,Gm1}
// ------------------------------------------------------------------------------------------------
template <typename T>
struct StreamIO;
template <>
struct StreamIO<double> {
    static __device__ __forceinline__ size_t index(int body, int c) { return (size_t)body * 4 + c; }
    static __device__ __forceinline__ double get(const double* s, int body, int c) { return s[index(body, c)]; }
    static __device__ __forceinline__ void put(double* s, int body, int c, double v) { s[index(body, c)] = v; }
};
template <>
struct StreamIO<float> {
    static __device__ __forceinline__ size_t index(int body, int c) {
        return (size_t)(body >> 1) * 8 + 2 * c + (body & 1);
    }
    static __device__ __forceinline__ float get(const float* s, int body, int c) { return s[index(body, c)]; }
    static __device__ __forceinline__ void put(float* s, int body, int c, float v) { s[index(body, c)] = v; }
};
// bytes of the stream per body (both layouts: 16 B in float32, 32 B in float64)
template <typename T>
constexpr int kStreamBytesPerBody = 4 * (int)sizeof(T);

// ------------------------------------------------------------------------------------------------
// float32: one tile of j PAIRS against kP bodies, two j per instruction through the packed f32x2 pipe
// (FADD2/FFMA2/FMUL2): 12 FMA-pipe lane-operations + 1 MUFU.RSQ per interaction.
// kGuard: the tile may hold one of this thread's own bodies.  The i == j term has r2 == eps2 and must contribute
// exactly 0 (the reference skips it, nbody.py:46).  dx = 0 is not enough: with eps = 1e-9, G*m*inv^3 overflows float32
// for G*m > 3.4e11 (any star) and inf * 0 = NaN.  Testing every pair costs two ALU instructions per lane -- measured:
// 72 % -> 63 % of the FP32 peak at N = 65,536 -- so the test (r2 > eps2 ? inv : 0) is compiled only into a second copy
// of the loop, taken for the few tiles that overlap the caller's own bodies.  eps == 0 needs nothing more: the same
// test removes r2 == 0.  Guarded and unguarded copies give the same bits for every other pair.
// ------------------------------------------------------------------------------------------------
template <int kP, bool kGuard>
__device__ __forceinline__ void f32_pairs(const float4* __restrict__ t, int n_pairs, const float (&xi)[kP],
                                          const float (&yi)[kP], const float (&zi)[kP], float2 (&ax)[kP],
                                          float2 (&ay)[kP], float2 (&az)[kP], float eps2) {
    const float2 e2 = make_float2(eps2, eps2);
    // j pairs in flight per thread: kP bodies already give kP independent chains per pair; one body per lane (the
    // group-per-CTA kernel at N <= 4,736) needs more pairs in flight to fill the pipe
    constexpr int kUnroll = kP == 1 ? 8 : 2;
#pragma unroll kUnroll
    for (int jp = 0; jp < n_pairs; ++jp) {
        const float4 A = t[2 * jp];      // x0 x1 y0 y1
        const float4 B = t[2 * jp + 1];  // z0 z1 gm0 gm1
        const float2 xj = make_float2(A.x, A.y), yj = make_float2(A.z, A.w);
        const float2 zj = make_float2(B.x, B.y), gj = make_float2(B.z, B.w);
#pragma unroll
        for (int k = 0; k < kP; ++k) {
            const float2 dx = __fadd2_rn(xj, make_float2(-xi[k], -xi[k]));
            const float2 dy = __fadd2_rn(yj, make_float2(-yi[k], -yi[k]));
            const float2 dz = __fadd2_rn(zj, make_float2(-zi[k], -zi[k]));
            float2 r2 = __ffma2_rn(dx, dx, e2);
            r2 = __ffma2_rn(dy, dy, r2);
            r2 = __ffma2_rn(dz, dz, r2);
            float2 inv;
            inv.x = rsqrt_approx(r2.x);
            inv.y = rsqrt_approx(r2.y);
            if (kGuard) {
                inv.x = (r2.x > eps2) ? inv.x : 0.f;
                inv.y = (r2.y > eps2) ? inv.y : 0.f;
            }
            const float2 inv2 = __fmul2_rn(inv, inv);
            float2 f = __fmul2_rn(gj, inv);
            f = __fmul2_rn(f, inv2);
            ax[k] = __ffma2_rn(f, dx, ax[k]);
            ay[k] = __ffma2_rn(f, dy, ay[k]);
            az[k] = __ffma2_rn(f, dz, az[k]);
        }
    }
}

// float64: one tile of j bodies against kP bodies: pair_f64(), 16 FP64-pipe operations + 1 MUFU.RSQ64H per interaction.
template <int kP, bool kZeroEps>
__device__ __forceinline__ void f64_bodies(const double2* __restrict__ t, int n_j, const double (&xi)[kP],
                                           const double (&yi)[kP], const double (&zi)[kP], double (&ax)[kP],
                                           double (&ay)[kP], double (&az)[kP], double eps2) {
    constexpr int kUnroll = kP == 1 ? 8 : 4;
#pragma unroll kUnroll
    for (int j = 0; j < n_j; ++j) {
        const double2 a = t[2 * j];      // x y
        const double2 b = t[2 * j + 1];  // z gm
        const double c1 = b.y * kC1OverC0;  // first-order build only (dead code otherwise): once per j
#pragma unroll
        for (int k = 0; k < kP; ++k)
            pair_f64<kZeroEps>(xi[k], yi[k], zi[k], a.x, a.y, b.x, b.y, c1, eps2, ax[k], ay[k], az[k]);
    }
}

// nb_persist.cu: all steps of a run in one cooperative launch (mid-size systems)
template <typename T>
int persist_run(T* stream_a, T* stream_b, T* vel, T* acc, int n, double dt, double softening, int n_steps,
                int save_interval, double* sp, double* sv, double* sa, T* partial, int* group_counter,
                unsigned* barrier, int* error, cudaStream_t st);

// nb_group.cu: one CTA per group of bodies, segment partials reduced in shared memory (whole systems that fit the
// SMs in one wave); group_step_kp() says whether a system qualifies (0: no) and with how many bodies per lane
int group_step_kp(int n, int n_seg, int sms, int is_f64);
template <typename T>
int group_step(const T* cur, T* next, T* vel, T* acc, int n, double dt, double softening, int mode, int flags,
               double* sp, double* sv, double* sa, int* error, int kP, cudaStream_t st, int B = 1, size_t snap_stride = 0);

}  // namespace nb
