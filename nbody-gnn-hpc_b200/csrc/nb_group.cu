// nb_group.cu -- K2s: the force pass + leapfrog of one step for whole systems of up to 64 bodies per SM (N <= 9,472 on
// 148 SMs), ONE CTA PER GROUP OF BODIES, no cross-CTA reduction (sm_100a).
//
// Replaces NBodySimulator.step, reference src/hpc/nbody.py:202-218, in the size range where K2 (nb_force.cu) spends
// most of a step on a chain of L2 round trips: there a body's j-segments are summed by DIFFERENT CTAs, so every step
// flushes the segment partials to global memory, fences, bumps the i-tile's arrival counter, and the last CTA reads
// them all back (~8 dependent round trips, 11 .. 17 us per step between N = 1,024 and 4,096, measured).
//
// Here a CTA owns a group of 32 x kP bodies for ALL j: its warp w is the warp-task (group, segment w) of K2's
// segment plan (nb_warptask.cuh: own TMA ring, K2's inner loops), the n_seg partials of a body meet in SHARED
// memory, and after one __syncthreads the first 32 x kP threads add them in ascending segment order and integrate
// the body -- closing kick, snapshot row, next opening kick, drift into the other stream buffer -- with the arithmetic
// of K2's epilogue: the same bits as K2 for every body, step after step (tested), hence still bit-identical to any
// i-slab / rank decomposition that uses K2.  One launch per step, chained by programmatic dependent launch like K2.
// kP (1 or 2) is the smallest for which the groups fit the SMs in one wave (N <= 64 x SMs = 9,472 on a B200); larger
// systems stay with K2, whose (i-tile x segment) grid fills the machine better.
#include "nb_warptask.cuh"

namespace nb {

constexpr int kGroupMaxWarps = 16;  // segments of nb_segment_plan for N <= 65,536

template <typename T>
struct GroupArgs {
    const T* cur;
    T* next;
    T* vel;
    T* acc;
    int n, n_pad, seg_len, n_seg;
    int mode;   // 1: accelerations only (nb_accel_*); 2: step (nb_step_*)
    int flags;  // NB_STEP_*
    T dt, half_dt, eps2;
    double* sp;
    double* sv;
    double* sa;
    int* error;
};

template <typename T, int kP, bool kZeroEps>
__global__ void __launch_bounds__(kGroupMaxWarps * 32, 1) group_step_kernel(const GroupArgs<T> g) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;  // n_warps == n_seg
    constexpr int kBodies = 32 * kP;
    char* ring = smem + (size_t)warp * kPStages * kPTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_warps * kPStages * kPTileBytes) + warp * kPStages;
    T* part = reinterpret_cast<T*>(smem + (size_t)n_warps * kPStages * (kPTileBytes + sizeof(uint64_t)));  // [seg][3][kBodies]
    if (lane == 0) {
        for (int s = 0; s < kPStages; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    __syncwarp();
    // programmatic dependent launch: nothing a predecessor wrote is touched before it has completed and flushed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0) asm volatile("fence.proxy.async;" ::: "memory");

    const int grp = blockIdx.x, b0 = grp * kBodies;
    const int j0 = warp * g.seg_len, j1 = min(j0 + g.seg_len, g.n_pad);
    uint32_t tiles_done = 0;
    const bool ok = WarpTask<T, kP, kZeroEps>::run(g.cur, g.n, grp, j0, j1, g.eps2, part + (size_t)warp * 3 * kBodies,
                                                   kBodies, b0, ring, bars, tiles_done, lane);
    if (!ok && lane == 0) atomicCAS(g.error, 0, NB_PERSIST_STALLED);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the successor's CTAs may be placed from here on
    __syncthreads();

    // the group's bodies: partials in ascending segment order, then K2's epilogue arithmetic (nbody.py:205-214)
    const int t = threadIdx.x;
    const int li = b0 + t;
    if (t >= kBodies || li >= g.n) return;
    T a[3] = {T(0), T(0), T(0)};
    for (int s = 0; s < n_warps; ++s) {
        const T* p = part + (size_t)s * 3 * kBodies;
#pragma unroll
        for (int c = 0; c < 3; ++c) a[c] += p[c * kBodies + t];
    }
    if (g.mode == 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) g.acc[(size_t)li * 3 + c] = a[c];
        return;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T x = g.cur[StreamIO<T>::index(li, c)];
        T v = g.vel[(size_t)li * 3 + c];
        v = mul_add_unfused(g.half_dt, a[c], v);  // closing kick, nbody.py:214
        if (g.flags & NB_STEP_SNAPSHOT) {          // get_state(), nbody.py:250-259
            if (g.sp) g.sp[(size_t)li * 3 + c] = (double)x;
            if (g.sv) g.sv[(size_t)li * 3 + c] = (double)v;
            if (g.sa) g.sa[(size_t)li * 3 + c] = (double)a[c];
        }
        if (g.flags & NB_STEP_CONTINUE) {
            v = mul_add_unfused(g.half_dt, a[c], v);  // next step's opening kick, nbody.py:205
            x = mul_add_unfused(g.dt, v, x);          // drift, nbody.py:208
            g.next[StreamIO<T>::index(li, c)] = x;
        }
        g.vel[(size_t)li * 3 + c] = v;
        g.acc[(size_t)li * 3 + c] = a[c];
    }
}

// The smallest kP in {1, 2} whose groups fit the SMs in one wave; 0: this system is K2's.  (kP = 4 would take systems
// up to 18,944 bodies, but its 96 .. 148 CTAs of 16 warps fill the machine worse than K2's grid does: measured 94.9 us
// against 77.8 us per step at N = 12,288 and 124 against 116 at 16,384, profiles/r02_midn_group.log.)
// In float64 kP = 2 is taken only when its groups fill at least 80 % of the SMs: below that K2 is faster (N = 5,000:
// 79 groups, 53.4 us against 41.6 us per step; from N ~ 7,600 up K2s wins again, 95 against 113 us at 9,472); in
// float32 K2s is at least as fast as K2 over the whole range (profiles/r02_midn_group.log).
int group_step_kp(int n, int n_seg, int sms, int is_f64) {
    if (n_seg < 1 || n_seg > kGroupMaxWarps) return 0;
    if (ceil_div(n, 32) <= sms) return 1;
    const int groups2 = ceil_div(n, 64);
    if (groups2 <= sms && (!is_f64 || groups2 * 10 >= sms * 8)) return 2;
    return 0;
}

template <typename T>
int group_step(const T* cur, T* next, T* vel, T* acc, int n, double dt, double softening, int mode, int flags,
               double* sp, double* sv, double* sa, int* error, int kP, cudaStream_t st) {
    GroupArgs<T> g;
    g.cur = cur; g.next = next; g.vel = vel; g.acc = acc;
    g.n = n; g.n_pad = nb_padded_bodies(n);
    nb_segment_plan(n, &g.seg_len, &g.n_seg);
    g.mode = mode; g.flags = flags;
    g.dt = (T)dt; g.half_dt = (T)(0.5 * dt);  // "0.5 * self.dt" is evaluated first, nbody.py:205
    g.eps2 = (T)(softening * softening);
    g.sp = sp; g.sv = sv; g.sa = sa; g.error = error;
    const bool zero = !(g.eps2 > T(0));
    void (*kern)(const GroupArgs<T>);
    if (kP == 2) kern = zero ? group_step_kernel<T, 2, true> : group_step_kernel<T, 2, false>;
    else kern = zero ? group_step_kernel<T, 1, true> : group_step_kernel<T, 1, false>;
    const size_t smem = (size_t)g.n_seg * (kPStages * (kPTileBytes + sizeof(uint64_t)) + 3 * 32 * kP * sizeof(T));
    static bool attr_set[2][2][2] = {};
    bool& done = attr_set[sizeof(T) == 8][kP - 1][zero];
    if (!done) {  // the permission for the largest shape, once per kernel
        const size_t smem_max = (size_t)kGroupMaxWarps * (kPStages * (kPTileBytes + sizeof(uint64_t)) + 3 * 32 * 2 * sizeof(double));
        NB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(n, 32 * kP));
    cfg.blockDim = dim3(g.n_seg * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, g));
    return check_launch("group step kernel");
}

template int group_step<float>(const float*, float*, float*, float*, int, double, double, int, int, double*, double*,
                               double*, int*, int, cudaStream_t);
template int group_step<double>(const double*, double*, double*, double*, int, double, double, int, int, double*,
                                double*, double*, int*, int, cudaStream_t);

}  // namespace nb
