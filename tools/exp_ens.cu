// exp_ens.cu -- development experiment (not part of the library): where does the ensemble kernel (K3) spend its
// time, and what do a statically-shaped integrate phase / direct snapshot stores / other CTA shapes buy?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/exp_ens tools/exp_ens.cu \
//              -Lnbody-gnn-hpc_b200/lib -lnbody_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../nbody-gnn-hpc_b200/lib'
// Run:    tools/exp_ens [B] [steps]
// Every variant is checked bit-for-bit against the library's nb_ensemble_f64 on the same inputs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../nbody-gnn-hpc_b200/csrc/nb_common.cuh"
using namespace nb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Args {
    double *x, *v, *a;
    const float* m32;
    int B, N;
    double dt, half_dt, eps2;
    int n_steps;
    double *ox, *ov, *oa;
    int n_snap;
    int* sm_slots;
    unsigned stagger_ns;
    long long* trace;  // per CTA: smid, t0, t1 (globaltimer ns), cyc_F, cyc_bar1, cyc_I, cyc_bar2
};

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }


// ---- seed experiments: does MUFU.RSQ64H cost an FP64-pipe slot? ----------------------------------------------
// kSeed 0: MUFU.RSQ64H (library).  1: fp32 MUFU.RSQ on an integer-converted operand (no F2F, no RSQ64H).
// 2: integer bit trick only (wrong numerics, timing only: no MUFU at all).
template <int kSeed>
__device__ __forceinline__ double seed_any(double r2) {
    if (kSeed == 0 || kSeed == 3) return rsqrt_seed(r2);
    const unsigned hi = (unsigned)__double2hiint(r2), lo = (unsigned)__double2loint(r2);
    if (kSeed == 1) {
        const unsigned fb = __funnelshift_l(lo, hi - 0x38000000u, 3);  // fp64 -> fp32 bits, truncated (normal range)
        const float y = rsqrt_approx(__uint_as_float(fb));
        const unsigned yb = __float_as_uint(y);
        return __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));
    }
    return __hiloint2double((int)(0x5fe6eb50u - (hi >> 1)), 0);
}
template <int kSeed>
__device__ __forceinline__ void pair_seed(double xi, double yi, double zi, double xj, double yj, double zj, double gmj,
                                          double eps2, double& ax, double& ay, double& az) {
    const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    const double r2b = r2;
    r2 = fma(dz, dz, r2);
    // kSeed 3: MUFU.RSQ64H only defines the high word; instead of zeroing the low word (one MOV per interaction) take
    // whatever is in the low word of the dying partial sum r2b: a 2^-20 relative perturbation of the seed.
    double y0;
    if (kSeed == 3) {
        y0 = r2b;  // dies here: its register pair takes the seed's high word in place
        asm("{\n\t.reg .b32 lo, hi, t;\n\t.reg .f64 y;\n\t"
            "rsqrt.approx.ftz.f64 y, %1;\n\t"
            "mov.b64 {t, hi}, y;\n\t"
            "mov.b64 {lo, t}, %0;\n\t"
            "mov.b64 %0, {lo, hi};\n\t}"
            : "+d"(y0)
            : "d"(r2));
    } else {
        y0 = seed_any<kSeed>(r2);
    }
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double g = gmj * y0;
    const double w = y2 * g;
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    const double f = fma(w, q, w);
    ax = fma(f, dx, ax); ay = fma(f, dy, ay); az = fma(f, dz, az);
}


// ---- stage-wise interleaving: kC independent interactions written stage by stage, so that ptxas keeps kC chains in
// flight instead of finishing one interaction before starting the next ----
template <int kC>
__device__ __forceinline__ void pairs_staged(const double (&xi)[kC], const double (&yi)[kC], const double (&zi)[kC],
                                             const double (&xj)[kC], const double (&yj)[kC], const double (&zj)[kC],
                                             const double (&gm)[kC], double eps2, double* (&ax)[kC], double* (&ay)[kC],
                                             double* (&az)[kC]) {
    double dx[kC], dy[kC], dz[kC], r2[kC], y0[kC], y2[kC], e[kC], g[kC], w[kC], p[kC], q[kC], f[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) { dx[c] = xj[c] - xi[c]; dy[c] = yj[c] - yi[c]; dz[c] = zj[c] - zi[c]; }
#pragma unroll
    for (int c = 0; c < kC; ++c) r2[c] = fma(dx[c], dx[c], eps2);
#pragma unroll
    for (int c = 0; c < kC; ++c) r2[c] = fma(dy[c], dy[c], r2[c]);
#pragma unroll
    for (int c = 0; c < kC; ++c) r2[c] = fma(dz[c], dz[c], r2[c]);
#pragma unroll
    for (int c = 0; c < kC; ++c) y0[c] = rsqrt_seed(r2[c]);
#pragma unroll
    for (int c = 0; c < kC; ++c) { y2[c] = y0[c] * y0[c]; g[c] = gm[c] * y0[c]; }
#pragma unroll
    for (int c = 0; c < kC; ++c) { e[c] = fma(-r2[c], y2[c], 1.0); w[c] = y2[c] * g[c]; }
#pragma unroll
    for (int c = 0; c < kC; ++c) p[c] = fma(1.875, e[c], 1.5);
#pragma unroll
    for (int c = 0; c < kC; ++c) q[c] = e[c] * p[c];
#pragma unroll
    for (int c = 0; c < kC; ++c) f[c] = fma(w[c], q[c], w[c]);
#pragma unroll
    for (int c = 0; c < kC; ++c) { *ax[c] = fma(f[c], dx[c], *ax[c]); *ay[c] = fma(f[c], dy[c], *ay[c]); *az[c] = fma(f[c], dz[c], *az[c]); }
}

// kN bodies, kRows x kParts force threads (two bodies per thread), kThreads per CTA.
// kDirect: snapshot rows are stored from the integrate phase straight to HBM (no shared stage, no flush pass).
// kRegs:   velocity / acceleration scalars live in registers of their owner thread.
template <int kN, int kRows, int kParts, int kThreads, int kMinBlocks, bool kDirect, bool kRegs, bool kTrace, int kUnroll, int kBPT, bool kForceOnly, int kSeed = 0, int kInter = 0>
__global__ void __launch_bounds__(kThreads, kMinBlocks) ens_v1(const Args g) {
    constexpr int n3 = 3 * kN;
    constexpr int kE = (n3 + kThreads - 1) / kThreads;  // integrator scalars per thread
    extern __shared__ __align__(16) char smem[];
    double4* pos = reinterpret_cast<double4*>(smem);
    double* pos_s = reinterpret_cast<double*>(smem);
    double* part = pos_s + 4 * kN;           // kParts x n3
    double* vel = part + kParts * n3;        // n3 (unused with kRegs)
    double* acc = vel + n3;                  // n3 (unused with kRegs)
    double* stage = acc + n3;                // 3 x n3 (unused with kDirect)

    const int tid = threadIdx.x;
    long long cF = 0, cB1 = 0, cI = 0, cB2 = 0, t0 = 0;
    unsigned smid = 0;
    if (kTrace && tid == 0) { asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); }

    if (g.sm_slots != nullptr && g.stagger_ns > 0) {
        if (tid == 0) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            const int arrival = atomicAdd(g.sm_slots + sm, 1);
            if (arrival & 1) {
                const long long c0 = clock64();
                const long long ticks = (long long)g.stagger_ns * 2;
                while (clock64() - c0 < ticks) __nanosleep(200);
            }
        }
        __syncthreads();
    }
    if (kTrace && tid == 0) t0 = gtime();

    const int q = tid / kRows;
    const int r = tid - q * kRows;
    const bool active = q < kParts;
    const int jb = active ? (q * kN) / kParts : 0;
    const int je = active ? ((q + 1) * kN) / kParts : 0;
    const double dt = g.dt, half_dt = g.half_dt, eps2 = g.eps2;

    // integrator ownership: scalar idx = tid + e * kThreads
    int poff[kE];
    bool own[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int idx = tid + e * kThreads;
        own[e] = idx < n3;
        const int i = idx / 3, c = idx - 3 * i;
        poff[e] = own[e] ? 4 * i + c : 0;
    }

    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        const size_t sbase = (size_t)b * n3;
        double vr[kE], ar[kE];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kE; ++e) {
            const int idx = tid + e * kThreads;
            if (own[e]) {
                pos_s[poff[e]] = g.x[sbase + idx];
                if (kRegs) { vr[e] = g.v[sbase + idx]; ar[e] = g.a[sbase + idx]; }
                else { vel[idx] = g.v[sbase + idx]; acc[idx] = g.a[sbase + idx]; }
            }
        }
        for (int i = tid; i < kN; i += kThreads) pos_s[4 * i + 3] = kG * (double)g.m32[i];
        __syncthreads();

        long pending = -1;
        for (int k = 0; k <= g.n_steps; ++k) {
            const bool do_close = k > 0;
            const bool do_open = k < g.n_steps;
            const long srow = k;  // save_interval 1, initial row written
            long long c0 = 0;
            if (kTrace && tid == 0) c0 = clock64();
            if (!kDirect && pending >= 0) {
                const size_t o = ((size_t)b * g.n_snap + (size_t)pending) * n3;
                for (int idx = tid; idx < n3; idx += kThreads) {
                    g.ox[o + idx] = stage[idx];
                    g.ov[o + idx] = stage[n3 + idx];
                    g.oa[o + idx] = stage[2 * n3 + idx];
                }
            }
            if (active) {
                double mx[kBPT], my[kBPT], mz[kBPT], ax[kBPT], ay[kBPT], az[kBPT];
#pragma unroll
                for (int m = 0; m < kBPT; ++m) {
                    const double4 me = pos[min(r + m * kRows, kN - 1)];
                    mx[m] = me.x; my[m] = me.y; mz[m] = me.z;
                    ax[m] = ay[m] = az[m] = 0.0;
                }
                if (kInter == 2 && kBPT == 2) {
#pragma unroll kUnroll
                    for (int j = jb; j < je; ++j) {
                        const double4 pj = pos[j];
                        const double xi_[2] = {mx[0], mx[1]}, yi_[2] = {my[0], my[1]}, zi_[2] = {mz[0], mz[1]};
                        const double xj_[2] = {pj.x, pj.x}, yj_[2] = {pj.y, pj.y}, zj_[2] = {pj.z, pj.z}, gm_[2] = {pj.w, pj.w};
                        double* axp[2] = {&ax[0], &ax[1]}; double* ayp[2] = {&ay[0], &ay[1]}; double* azp[2] = {&az[0], &az[1]};
                        pairs_staged<2>(xi_, yi_, zi_, xj_, yj_, zj_, gm_, eps2, axp, ayp, azp);
                    }
                } else if (kInter == 4 && kBPT == 2) {
#pragma unroll kUnroll
                    for (int j = jb; j < je; j += 2) {
                        const double4 pa = pos[j], pb = pos[j + 1];
                        const double xi_[4] = {mx[0], mx[1], mx[0], mx[1]}, yi_[4] = {my[0], my[1], my[0], my[1]}, zi_[4] = {mz[0], mz[1], mz[0], mz[1]};
                        const double xj_[4] = {pa.x, pa.x, pb.x, pb.x}, yj_[4] = {pa.y, pa.y, pb.y, pb.y}, zj_[4] = {pa.z, pa.z, pb.z, pb.z}, gm_[4] = {pa.w, pa.w, pb.w, pb.w};
                        double* axp[4] = {&ax[0], &ax[1], &ax[0], &ax[1]}; double* ayp[4] = {&ay[0], &ay[1], &ay[0], &ay[1]}; double* azp[4] = {&az[0], &az[1], &az[0], &az[1]};
                        // chains 0/2 and 1/3 share accumulators: add the pair (0,1) first, then (2,3)
                        double* a01x[2] = {axp[0], axp[1]}; (void)a01x;
                        pairs_staged<4>(xi_, yi_, zi_, xj_, yj_, zj_, gm_, eps2, axp, ayp, azp);
                    }
                } else {
#pragma unroll kUnroll
                for (int j = jb; j < je; ++j) {
                    const double4 pj = pos[j];
#pragma unroll
                    for (int m = 0; m < kBPT; ++m)
                        pair_seed<kSeed>(mx[m], my[m], mz[m], pj.x, pj.y, pj.z, pj.w, eps2, ax[m], ay[m], az[m]);
                }
                }
                double* pa = part + q * n3;
#pragma unroll
                for (int m = 0; m < kBPT; ++m) {
                    const int im = r + m * kRows;
                    if (im < kN) { pa[3 * im + 0] = ax[m]; pa[3 * im + 1] = ay[m]; pa[3 * im + 2] = az[m]; }
                }
            }
            long long c1 = 0;
            if (kTrace && tid == 0) { c1 = clock64(); cF += c1 - c0; }
            __syncthreads();
            long long c2 = 0;
            if (kTrace && tid == 0) { c2 = clock64(); cB1 += c2 - c1; }
            const size_t o = ((size_t)b * g.n_snap + (size_t)srow) * n3;
            if (!kForceOnly)
#pragma unroll
            for (int e = 0; e < kE; ++e) {
                const int idx = tid + e * kThreads;
                if (own[e]) {
                    double a = part[idx];
#pragma unroll
                    for (int p = 1; p < kParts; ++p) a += part[p * n3 + idx];
                    double v = kRegs ? vr[e] : vel[idx];
                    double x = pos_s[poff[e]];
                    if (do_close) v = mul_add_unfused(half_dt, a, v);
                    if (kDirect) {
                        g.ox[o + idx] = x; g.ov[o + idx] = v; g.oa[o + idx] = a;
                    } else {
                        stage[idx] = x; stage[n3 + idx] = v; stage[2 * n3 + idx] = a;
                    }
                    if (do_open) {
                        v = mul_add_unfused(half_dt, a, v);
                        x = mul_add_unfused(dt, v, x);
                        pos_s[poff[e]] = x;
                    }
                    if (kRegs) { vr[e] = v; ar[e] = a; }
                    else { vel[idx] = v; acc[idx] = a; }
                }
            }
            pending = srow;
            long long c3 = 0;
            if (kTrace && tid == 0) { c3 = clock64(); cI += c3 - c2; }
            __syncthreads();
            if (kTrace && tid == 0) cB2 += clock64() - c3;
        }
        if (!kDirect && pending >= 0) {
            const size_t o = ((size_t)b * g.n_snap + (size_t)pending) * n3;
            for (int idx = tid; idx < n3; idx += kThreads) {
                g.ox[o + idx] = stage[idx];
                g.ov[o + idx] = stage[n3 + idx];
                g.oa[o + idx] = stage[2 * n3 + idx];
            }
        }
#pragma unroll
        for (int e = 0; e < kE; ++e) {
            const int idx = tid + e * kThreads;
            if (own[e]) {
                g.x[sbase + idx] = pos_s[poff[e]];
                g.v[sbase + idx] = kRegs ? vr[e] : vel[idx];
                g.a[sbase + idx] = kRegs ? ar[e] : acc[idx];
            }
        }
    }
    if (kTrace && tid == 0) {
        long long* t = g.trace + (size_t)blockIdx.x * 8;
        t[0] = smid; t[1] = t0; t[2] = gtime(); t[3] = cF; t[4] = cB1; t[5] = cI; t[6] = cB2;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Symmetric variant: every unordered pair is evaluated once (Newton's third law), 20 FP64 operations per pair =
// 10 per ordered interaction instead of 16.  Bodies are grouped in blocks of 8; thread t owns one block pair
// (I, J = I + k mod nb), k = 1 .. (nb-1)/2, i.e. an 8x8 tile of pairs: the 8 bodies i live in registers with their
// accumulators, the 8 bodies j are streamed; both sides' partial sums go to per-(body, slot) words in shared memory,
// where slot names the partner block, so that the integrate phase adds them in one fixed order.  nb extra threads
// own the diagonal tiles (one-sided, 16 operations per ordered pair).
// ---------------------------------------------------------------------------------------------------------------
template <int kN, int kThreads, bool kTrace, int kUnrollJ>
__global__ void __launch_bounds__(kThreads, 1) ens_sym(const Args g) {
    constexpr int nb = kN / 8;               // blocks (kN % 8 == 0, nb odd in this experiment)
    constexpr int kHalf = (nb - 1) / 2;      // ring offsets k = 1 .. kHalf
    constexpr int nOff = nb * kHalf;         // off-diagonal tiles
    constexpr int kSlots = 2 * kHalf + 1;    // partials per body
    constexpr int kSlotPad = kSlots + (kSlots & 1);  // even: 16-byte aligned rows
    constexpr int n3 = 3 * kN;
    constexpr int kE = (n3 + kThreads - 1) / kThreads;
    constexpr int kBlkStride = 34;           // doubles per block of 8 bodies: 8 x {x,y,z,gm} + 16 bytes of padding
    extern __shared__ __align__(16) char smem[];
    double* pos_s = reinterpret_cast<double*>(smem);          // nb x 34
    double* part = pos_s + nb * kBlkStride;                   // [c][jj][blk][slot]
    double* vel = part + (size_t)3 * 8 * nb * kSlotPad;
    double* acc = vel + n3;

    const int tid = threadIdx.x;
    long long cF = 0, cB1 = 0, cI = 0, cB2 = 0, t0 = 0;
    unsigned smid = 0;
    if (kTrace && tid == 0) { asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); t0 = gtime(); }
    const double dt = g.dt, half_dt = g.half_dt, eps2 = g.eps2;

    // tile of this thread
    const bool off = tid < nOff;
    const bool diag = !off && tid < nOff + nb;
    int I = 0, J = 0, slotI = 0, slotJ = 0;
    if (off) { const int k = tid / nb + 1; I = tid - (k - 1) * nb; J = I + k; if (J >= nb) J -= nb; slotI = k - 1; slotJ = kHalf + k - 1; }
    if (diag) { I = J = tid - nOff; slotI = 2 * kHalf; }

    int poff[kE], qoff[kE];
    bool own[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int idx = tid + e * kThreads;
        own[e] = idx < n3;
        const int i = own[e] ? idx / 3 : 0, c = own[e] ? idx - 3 * i : 0;
        poff[e] = (i >> 3) * kBlkStride + (i & 7) * 4 + c;
        qoff[e] = ((c * 8 + (i & 7)) * nb + (i >> 3)) * kSlotPad;
    }

    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        const size_t sbase = (size_t)b * n3;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kE; ++e) {
            const int idx = tid + e * kThreads;
            if (own[e]) { pos_s[poff[e]] = g.x[sbase + idx]; vel[idx] = g.v[sbase + idx]; acc[idx] = g.a[sbase + idx]; }
        }
        for (int i = tid; i < kN; i += kThreads) pos_s[(i >> 3) * kBlkStride + (i & 7) * 4 + 3] = kG * (double)g.m32[i];
        __syncthreads();

        for (int k = 0; k <= g.n_steps; ++k) {
            const bool do_force = k > 0;  // a_0 is given
            const bool do_close = k > 0;
            const bool do_open = k < g.n_steps;
            long long c0 = 0;
            if (kTrace && tid == 0) c0 = clock64();
            if (do_force && (off || diag)) {
                double xi[8], yi[8], zi[8], gi[8], ax[8], ay[8], az[8];
                const double2* bi = reinterpret_cast<const double2*>(pos_s + I * kBlkStride);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const double2 u = bi[2 * m], w = bi[2 * m + 1];
                    xi[m] = u.x; yi[m] = u.y; zi[m] = w.x; gi[m] = w.y;
                    ax[m] = ay[m] = az[m] = 0.0;
                }
                const double2* bj = reinterpret_cast<const double2*>(pos_s + J * kBlkStride);
                if (off) {
                    double* pj = part + (size_t)J * kSlotPad + slotJ;
#pragma unroll kUnrollJ
                    for (int jj = 0; jj < 8; ++jj) {
                        const double2 u = bj[2 * jj], w = bj[2 * jj + 1];
                        double bx = 0.0, by = 0.0, bz = 0.0;
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const double dx = u.x - xi[m], dy = u.y - yi[m], dz = w.x - zi[m];
                            double r2 = fma(dx, dx, eps2);
                            r2 = fma(dy, dy, r2);
                            r2 = fma(dz, dz, r2);
                            const double y0 = rsqrt_seed(r2);
                            const double y2 = y0 * y0;
                            const double e = fma(-r2, y2, 1.0);
                            const double y3 = y2 * y0;
                            const double p = fma(1.875, e, 1.5);
                            const double q = e * p;
                            const double s3 = fma(y3, q, y3);  // r2^(-3/2)
                            const double fi = w.y * s3;         // G m_j / r^3
                            const double fj = gi[m] * s3;       // G m_i / r^3
                            ax[m] = fma(fi, dx, ax[m]); ay[m] = fma(fi, dy, ay[m]); az[m] = fma(fi, dz, az[m]);
                            bx = fma(-fj, dx, bx); by = fma(-fj, dy, by); bz = fma(-fj, dz, bz);
                        }
                        pj[(size_t)(0 * 8 + jj) * nb * kSlotPad] = bx;
                        pj[(size_t)(1 * 8 + jj) * nb * kSlotPad] = by;
                        pj[(size_t)(2 * 8 + jj) * nb * kSlotPad] = bz;
                    }
                } else {
#pragma unroll kUnrollJ
                    for (int jj = 0; jj < 8; ++jj) {
                        const double2 u = bj[2 * jj], w = bj[2 * jj + 1];
#pragma unroll
                        for (int m = 0; m < 8; ++m)
                            pair_f64<false>(xi[m], yi[m], zi[m], u.x, u.y, w.x, w.y, eps2, ax[m], ay[m], az[m]);
                    }
                }
                double* pi = part + (size_t)I * kSlotPad + slotI;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    pi[(size_t)(0 * 8 + m) * nb * kSlotPad] = ax[m];
                    pi[(size_t)(1 * 8 + m) * nb * kSlotPad] = ay[m];
                    pi[(size_t)(2 * 8 + m) * nb * kSlotPad] = az[m];
                }
            }
            long long c1 = 0;
            if (kTrace && tid == 0) { c1 = clock64(); cF += c1 - c0; }
            __syncthreads();
            long long c2 = 0;
            if (kTrace && tid == 0) { c2 = clock64(); cB1 += c2 - c1; }
            const size_t o = ((size_t)b * g.n_snap + (size_t)k) * n3;
#pragma unroll
            for (int e = 0; e < kE; ++e) {
                const int idx = tid + e * kThreads;
                if (own[e]) {
                    double a = acc[idx];
                    if (do_force) {
                        const double2* pr = reinterpret_cast<const double2*>(part + qoff[e]);
                        double2 t2 = pr[0];
                        a = t2.x + t2.y;
#pragma unroll
                        for (int sidx = 1; sidx < kSlotPad / 2; ++sidx) {
                            t2 = pr[sidx];
                            a += t2.x;
                            if (2 * sidx + 1 < kSlots) a += t2.y;
                        }
                        acc[idx] = a;
                    }
                    double v = vel[idx];
                    double x = pos_s[poff[e]];
                    if (do_close) v = mul_add_unfused(half_dt, a, v);
                    g.ox[o + idx] = x; g.ov[o + idx] = v; g.oa[o + idx] = a;
                    if (do_open) {
                        v = mul_add_unfused(half_dt, a, v);
                        x = mul_add_unfused(dt, v, x);
                        pos_s[poff[e]] = x;
                    }
                    vel[idx] = v;
                }
            }
            long long c3 = 0;
            if (kTrace && tid == 0) { c3 = clock64(); cI += c3 - c2; }
            __syncthreads();
            if (kTrace && tid == 0) cB2 += clock64() - c3;
        }
#pragma unroll
        for (int e = 0; e < kE; ++e) {
            const int idx = tid + e * kThreads;
            if (own[e]) { g.x[sbase + idx] = pos_s[poff[e]]; g.v[sbase + idx] = vel[idx]; g.a[sbase + idx] = acc[idx]; }
        }
    }
    if (kTrace && tid == 0) {
        long long* t = g.trace + (size_t)blockIdx.x * 8;
        t[0] = smid; t[1] = t0; t[2] = gtime(); t[3] = cF; t[4] = cB1; t[5] = cI; t[6] = cB2;
    }
}

struct Host {
    int B, N, steps, n_snap;
    std::vector<double> x, v, a;
    std::vector<float> m;
    double *dx, *dv, *da, *ox, *ov, *oa, *rx, *rv, *ra;  // r*: reference outputs from the library
    float* dm;
    void* ws; size_t ws_bytes;
    int* slots;
    long long* trace;
};

static void reset_state(Host& h) {
    CK(cudaMemcpy(h.dx, h.x.data(), h.x.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h.dv, h.v.data(), h.v.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h.da, h.a.data(), h.a.size() * 8, cudaMemcpyHostToDevice));
}

template <class F>
static float time_ms(Host& h, int reps, F&& launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps + 1; ++r) {
        reset_state(h);
        CK(cudaMemset(h.slots, 0, 4096));
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) { best = std::min(best, ms); sum += ms; }
    }
    printf("  best %.3f ms  mean %.3f ms", best, sum / reps);
    return best;
}

// largest relative difference (per-body vector norm) of snapshot row `row` between the experiment and the library
static double row_diff(Host& h, double* mine, double* ref, int row) {
    const size_t n3 = (size_t)h.N * 3;
    std::vector<double> p(n3), q(n3);
    double worst = 0;
    for (int b = 0; b < h.B; b += 37) {
        const size_t o = ((size_t)b * h.n_snap + row) * n3;
        CK(cudaMemcpy(p.data(), mine + o, n3 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(q.data(), ref + o, n3 * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < h.N; ++i) {
            double d = 0, n = 0;
            for (int c = 0; c < 3; ++c) { const double u = p[3 * i + c] - q[3 * i + c]; d += u * u; n += q[3 * i + c] * q[3 * i + c]; }
            worst = std::max(worst, sqrt(d / n));
        }
    }
    return worst;
}

static bool same(Host& h, const char* what) {
    const size_t n = (size_t)h.B * h.n_snap * h.N * 3;
    std::vector<double> p(n), q(n);
    bool ok = true;
    double* o[3] = {h.ox, h.ov, h.oa};
    double* r[3] = {h.rx, h.rv, h.ra};
    for (int k = 0; k < 3; ++k) {
        CK(cudaMemcpy(p.data(), o[k], n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(q.data(), r[k], n * 8, cudaMemcpyDeviceToHost));
        if (memcmp(p.data(), q.data(), n * 8) != 0) ok = false;
    }
    printf("  %s: %s\n", what, ok ? "bit-identical to the library" : "DIFFERS from the library");
    return ok;
}

template <int kN, int kRows, int kParts, int kThreads, int kMinBlocks, bool kDirect, bool kRegs, int kUnroll, int kBPT = 2, bool kForceOnly = false, int kSeed = 0, int kInter = 0>
static void run_variant(Host& h, const char* name, bool stagger, bool trace) {
    constexpr int n3 = 3 * kN;
    const size_t smem = (size_t)kN * 32 + (size_t)(kParts + 2 + 3) * n3 * 8;
    Args g{};
    g.x = h.dx; g.v = h.dv; g.a = h.da; g.m32 = h.dm; g.B = h.B; g.N = kN;
    g.dt = 1e-3; g.half_dt = 0.5e-3; g.eps2 = 1e-18; g.n_steps = h.steps;
    g.ox = h.ox; g.ov = h.ov; g.oa = h.oa; g.n_snap = h.n_snap;
    g.sm_slots = stagger ? h.slots : nullptr;
    g.stagger_ns = (unsigned)((double)kN * kN * 16.0 / (64.0 * 1.9));
    g.trace = h.trace;
    auto kern = ens_v1<kN, kRows, kParts, kThreads, kMinBlocks, kDirect, kRegs, false, kUnroll, kBPT, kForceOnly, kSeed, kInter>;
    auto kern_t = ens_v1<kN, kRows, kParts, kThreads, kMinBlocks, kDirect, kRegs, true, kUnroll, kBPT, kForceOnly, kSeed, kInter>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    const int grid = std::min(h.B, per_sm * 148);
    printf("%-44s regs %3d smem %6zu per_sm %d grid %d stagger %d\n", name, fa.numRegs, smem, per_sm, grid, (int)stagger);
    CK(cudaMemset(h.ox, 0, (size_t)h.B * h.n_snap * n3 * 8));
    time_ms(h, 5, [&] { kern<<<grid, kThreads, smem>>>(g); });
    printf("\n");
    if (!same(h, name)) printf("  vs library: acc row 1 rel %.2e, pos row 2 rel %.2e, pos row 10 rel %.2e\n", row_diff(h, h.oa, h.ra, 1), row_diff(h, h.ox, h.rx, 2), row_diff(h, h.ox, h.rx, 10));
    if (trace) {
        reset_state(h);
        CK(cudaMemset(h.slots, 0, 4096));
        kern_t<<<grid, kThreads, smem>>>(g);
        CK(cudaDeviceSynchronize());
        std::vector<long long> t((size_t)grid * 8);
        CK(cudaMemcpy(t.data(), h.trace, t.size() * 8, cudaMemcpyDeviceToHost));
        long long tmin = t[1], tmax = 0;
        for (int c = 0; c < grid; ++c) { tmin = std::min(tmin, t[c * 8 + 1]); tmax = std::max(tmax, t[c * 8 + 2]); }
        std::vector<double> dur(grid), endt(grid);
        double sF = 0, sB1 = 0, sI = 0, sB2 = 0;
        for (int c = 0; c < grid; ++c) {
            dur[c] = (t[c * 8 + 2] - t[c * 8 + 1]) * 1e-6;
            endt[c] = (t[c * 8 + 2] - tmin) * 1e-6;
            sF += t[c * 8 + 3]; sB1 += t[c * 8 + 4]; sI += t[c * 8 + 5]; sB2 += t[c * 8 + 6];
        }
        std::vector<double> e2 = endt; std::sort(e2.begin(), e2.end());
        const double per = 1.0 / grid / (h.steps + 1);
        printf("  trace: span %.3f ms; CTA end times min %.3f p25 %.3f median %.3f p75 %.3f max %.3f ms\n",
               (tmax - tmin) * 1e-6, e2.front(), e2[grid / 4], e2[grid / 2], e2[3 * grid / 4], e2.back());
        printf("  thread-0 cycles per step: force %.0f  barrier1 %.0f  integrate %.0f  barrier2 %.0f  (sum %.0f)\n",
               sF * per, sB1 * per, sI * per, sB2 * per, (sF + sB1 + sI + sB2) * per);
        // co-resident pairs: difference of end times on the same SM
        std::vector<std::vector<int>> by_sm(256);
        for (int c = 0; c < grid; ++c) by_sm[t[c * 8] & 255].push_back(c);
        double dmax = 0, dsum = 0; int np = 0, low_first = 0, early_first = 0, pair_148 = 0;
        for (auto& v : by_sm) if (v.size() == 2) {
            const int c0 = v[0], c1 = v[1];  // c0 < c1
            const double d = fabs(endt[c0] - endt[c1]); dmax = std::max(dmax, d); dsum += d; ++np;
            const int winner = endt[c0] < endt[c1] ? c0 : c1;
            low_first += winner == c0;
            early_first += (t[winner * 8 + 1] <= t[(winner == c0 ? c1 : c0) * 8 + 1]);
            pair_148 += (c1 - c0 == 148);
        }
        if (np) printf("  co-resident CTA pairs: %d, end-time difference mean %.3f ms max %.3f ms; lower blockIdx finishes first in %d, "
                       "earlier starter first in %d, pairs (m, m+148): %d\n", np, dsum / np, dmax, low_first, early_first, pair_148);
    }
}


template <int kN, int kThreads, int kUnrollJ>
static void run_sym(Host& h, const char* name) {
    constexpr int nb = kN / 8, kHalf = (nb - 1) / 2, kSlots = 2 * kHalf + 1, kSlotPad = kSlots + (kSlots & 1);
    const size_t smem = ((size_t)nb * 34 + (size_t)3 * 8 * nb * kSlotPad + 6 * kN) * 8;
    Args g{};
    g.x = h.dx; g.v = h.dv; g.a = h.da; g.m32 = h.dm; g.B = h.B; g.N = kN;
    g.dt = 1e-3; g.half_dt = 0.5e-3; g.eps2 = 1e-18; g.n_steps = h.steps;
    g.ox = h.ox; g.ov = h.ov; g.oa = h.oa; g.n_snap = h.n_snap;
    g.trace = h.trace;
    auto kern = ens_sym<kN, kThreads, false, kUnrollJ>;
    auto kern_t = ens_sym<kN, kThreads, true, kUnrollJ>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    const int grid = std::min(h.B, per_sm * 148);
    printf("%-44s regs %3d smem %6zu per_sm %d grid %d\n", name, fa.numRegs, smem, per_sm, grid);
    CK(cudaMemset(h.ox, 0, (size_t)h.B * h.n_snap * kN * 3 * 8));
    time_ms(h, 5, [&] { kern<<<grid, kThreads, smem>>>(g); });
    printf("\n  vs library: acc row 1 rel %.2e, pos row 2 rel %.2e, pos row 10 rel %.2e\n", row_diff(h, h.oa, h.ra, 1),
           row_diff(h, h.ox, h.rx, 2), row_diff(h, h.ox, h.rx, 10));
    reset_state(h);
    kern_t<<<grid, kThreads, smem>>>(g);
    CK(cudaDeviceSynchronize());
    std::vector<long long> t((size_t)grid * 8);
    CK(cudaMemcpy(t.data(), h.trace, t.size() * 8, cudaMemcpyDeviceToHost));
    double sF = 0, sB1 = 0, sI = 0, sB2 = 0;
    for (int c = 0; c < grid; ++c) { sF += t[c * 8 + 3]; sB1 += t[c * 8 + 4]; sI += t[c * 8 + 5]; sB2 += t[c * 8 + 6]; }
    const double per = 1.0 / h.B / (h.steps + 1);
    printf("  thread-0 cycles per system-step: force %.0f  barrier1 %.0f  integrate %.0f  barrier2 %.0f  (sum %.0f)\n",
           sF * per, sB1 * per, sI * per, sB2 * per, (sF + sB1 + sI + sB2) * per);
}

int main(int argc, char** argv) {
    Host h;
    h.B = argc > 1 ? atoi(argv[1]) : 296;
    h.steps = argc > 2 ? atoi(argv[2]) : 400;
    h.N = 200;
    h.n_snap = h.steps + 1;
    const size_t ns = (size_t)h.B * h.N * 3;
    h.x.resize(ns); h.v.resize(ns); h.a.assign(ns, 0.0); h.m.resize(h.N);
    srand(42);
    for (size_t i = 0; i < ns; ++i) { h.x[i] = (rand() / (double)RAND_MAX - 0.5) * 10.0; h.v[i] = (rand() / (double)RAND_MAX - 0.5); }
    for (int i = 0; i < h.N; ++i) h.m[i] = (float)(1e10 + (1e12 - 1e10) * (rand() / (double)RAND_MAX));
    const size_t on = (size_t)h.B * h.n_snap * h.N * 3 * 8;
    CK(cudaMalloc(&h.dx, ns * 8)); CK(cudaMalloc(&h.dv, ns * 8)); CK(cudaMalloc(&h.da, ns * 8));
    CK(cudaMalloc(&h.dm, h.N * 4));
    CK(cudaMalloc(&h.ox, on)); CK(cudaMalloc(&h.ov, on)); CK(cudaMalloc(&h.oa, on));
    CK(cudaMalloc(&h.rx, on)); CK(cudaMalloc(&h.rv, on)); CK(cudaMalloc(&h.ra, on));
    CK(cudaMalloc(&h.slots, 4096)); CK(cudaMalloc(&h.trace, 8 * 8 * 1024));
    CK(cudaMemcpy(h.dm, h.m.data(), h.N * 4, cudaMemcpyHostToDevice));
    h.ws_bytes = nb_ensemble_workspace_bytes(h.B);
    CK(cudaMalloc(&h.ws, h.ws_bytes));

    // a0 for every system through the library (compute_a0 with 0 steps), so that the variants start from (x, v, a)_0
    reset_state(h);
    if (nb_ensemble_f64(h.dx, h.dv, h.da, h.dm, 1, 0, h.B, h.N, 1e-3, 1e-9, 0, 1, 1, 0, nullptr, nullptr, nullptr, 0, 0, h.ws,
                        h.ws_bytes, nullptr)) { printf("library call failed: %s\n", nb_last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.a.data(), h.da, ns * 8, cudaMemcpyDeviceToHost));

    printf("B=%d N=%d steps=%d\n", h.B, h.N, h.steps);
    printf("%-44s\n", "library nb_ensemble_f64");
    time_ms(h, 5, [&] {
        nb_ensemble_f64(h.dx, h.dv, h.da, h.dm, 1, 0, h.B, h.N, 1e-3, 1e-9, h.steps, 1, 0, 1, h.rx, h.rv, h.ra, h.n_snap, 0,
                        h.ws, h.ws_bytes, nullptr);
    });
    printf("\n");

    printf("---- force phase only ----\n");
    run_variant<200, 100, 5, 512, 1, true, true, 4, 2, true, 0, 0>(h, "F-only 1 CTA/SM, RSQ64H + zeroed low word", false, false);
    run_variant<200, 100, 5, 512, 1, true, true, 4, 2, true, 3, 0>(h, "F-only 1 CTA/SM, RSQ64H + stale low word", false, false);
    run_variant<200, 100, 5, 512, 2, true, true, 4, 2, true, 0, 0>(h, "F-only 2 CTA/SM, RSQ64H + zeroed low word", false, false);
    run_variant<200, 100, 5, 512, 2, true, true, 4, 2, true, 3, 0>(h, "F-only 2 CTA/SM, RSQ64H + stale low word", false, false);
    run_variant<200, 100, 5, 512, 1, true, true, 4, 2, false, 3, 0>(h, "full 1 CTA/SM, RSQ64H + stale low word", false, false);
    return 0;
}
