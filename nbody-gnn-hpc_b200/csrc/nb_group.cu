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
    int* error;  // may be null (batched launches)
    // batched launches (blockIdx.y = system): element strides between consecutive systems; 0 for a single system
    size_t stream_stride, va_stride, snap_stride;
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
    // this CTA's system (batched launches: independent systems side by side in the grid's y dimension)
    const size_t sys = blockIdx.y;
    const T* cur = g.cur + sys * g.stream_stride;
    T* next = g.next ? g.next + sys * g.stream_stride : nullptr;
    T* vel = g.vel ? g.vel + sys * g.va_stride : nullptr;
    T* acc = g.acc + sys * g.va_stride;
    double* sp = g.sp ? g.sp + sys * g.snap_stride : nullptr;
    double* sv = g.sv ? g.sv + sys * g.snap_stride : nullptr;
    double* sa = g.sa ? g.sa + sys * g.snap_stride : nullptr;
    uint32_t tiles_done = 0;
    const bool ok = WarpTask<T, kP, kZeroEps>::run(cur, g.n, grp, j0, j1, g.eps2, part + (size_t)warp * 3 * kBodies,
                                                   kBodies, b0, ring, bars, tiles_done, lane);
    if (!ok) {  // a tile copy never completed: never a silent wrong answer
        if (g.error == nullptr) __trap();  // batched launches have no error word: the launch fails, the host sees it
        if (lane == 0) atomicCAS(g.error, 0, NB_PERSIST_STALLED);
    }
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
        for (int c = 0; c < 3; ++c) acc[(size_t)li * 3 + c] = a[c];
        return;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T x = cur[StreamIO<T>::index(li, c)];
        T v = vel[(size_t)li * 3 + c];
        v = mul_add_unfused(g.half_dt, a[c], v);  // closing kick, nbody.py:214
        if (g.flags & NB_STEP_SNAPSHOT) {          // get_state(), nbody.py:250-259
            if (sp) sp[(size_t)li * 3 + c] = (double)x;
            if (sv) sv[(size_t)li * 3 + c] = (double)v;
            if (sa) sa[(size_t)li * 3 + c] = (double)a[c];
        }
        if (g.flags & NB_STEP_CONTINUE) {
            v = mul_add_unfused(g.half_dt, a[c], v);  // next step's opening kick, nbody.py:205
            x = mul_add_unfused(g.dt, v, x);          // drift, nbody.py:208
            next[StreamIO<T>::index(li, c)] = x;
        }
        vel[(size_t)li * 3 + c] = v;
        acc[(size_t)li * 3 + c] = a[c];
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
               double* sp, double* sv, double* sa, int* error, int kP, cudaStream_t st, int B, size_t snap_stride) {
    GroupArgs<T> g;
    g.cur = cur; g.next = next; g.vel = vel; g.acc = acc;
    g.stream_stride = B > 1 ? (size_t)nb_padded_bodies(n) * 4 : 0;
    g.va_stride = B > 1 ? (size_t)n * 3 : 0;
    g.snap_stride = B > 1 ? snap_stride : 0;
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
    cfg.gridDim = dim3(ceil_div(n, 32 * kP), B);
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
                               double*, int*, int, cudaStream_t, int, size_t);
template int group_step<double>(const double*, double*, double*, double*, int, double, double, int, int, double*,
                                double*, double*, int*, int, cudaStream_t, int, size_t);

// ------------------------------------------------------------------------------------------------------------------
// B independent mid-size systems side by side (ensembles of systems too large for the one-CTA-per-system kernel K3)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBatchedMaxBodies = 16384;  // n_seg <= 16 warps per CTA

// kP of a batched launch: waves do not matter here (B x groups CTAs), the per-CTA efficiency does
static int batched_kp(int n) { return n > 2048 ? 2 : 1; }

template <typename T>
__global__ void __launch_bounds__(256)
batched_kick_drift_kernel(const T* __restrict__ cur, T* __restrict__ next, T* __restrict__ vel, const T* __restrict__ acc,
                          int n, size_t stream_stride, T dt, T half_dt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t b = blockIdx.y;
    cur += b * stream_stride; next += b * stream_stride; vel += b * (size_t)n * 3; acc += b * (size_t)n * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T v = vel[(size_t)i * 3 + c];
        v = mul_add_unfused(half_dt, acc[(size_t)i * 3 + c], v);                     // nbody.py:205
        const T x = mul_add_unfused(dt, v, cur[StreamIO<T>::index(i, c)]);          // nbody.py:208
        next[StreamIO<T>::index(i, c)] = x;
        vel[(size_t)i * 3 + c] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
batched_snapshot0_kernel(const T* __restrict__ stream, const T* __restrict__ vel, const T* __restrict__ acc, int n,
                         size_t stream_stride, size_t snap_stride, double* __restrict__ sp, double* __restrict__ sv,
                         double* __restrict__ sa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t b = blockIdx.y;
    stream += b * stream_stride; vel += b * (size_t)n * 3; acc += b * (size_t)n * 3;
    sp += b * snap_stride; sv += b * snap_stride; sa += b * snap_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        sp[(size_t)i * 3 + c] = (double)stream[StreamIO<T>::index(i, c)];
        sv[(size_t)i * 3 + c] = (double)vel[(size_t)i * 3 + c];
        sa[(size_t)i * 3 + c] = (double)acc[(size_t)i * 3 + c];
    }
}

template <typename T>
static int accel_batched(const T* streams, int B, int n, double softening, T* acc, cudaStream_t st) {
    NB_REQUIRE(streams && acc && B > 0 && n > 0 && n <= kBatchedMaxBodies, "nb_accel_batched: need pointers, B > 0 and 0 < n <= %d",
               kBatchedMaxBodies);
    return group_step<T>(streams, nullptr, nullptr, acc, n, 0.0, softening, /*mode=*/1, 0, nullptr, nullptr, nullptr,
                         nullptr, batched_kp(n), st, B, 0);
}

// The loop of NBodySimulator.run (nbody.py:232-248) for B systems at once: (x_0, v_0, a_0) in, snapshots of every
// save_interval-th step out, one batched launch per step.
template <typename T>
static int run_batched(T* sa_, T* sb_, T* vel, T* acc, int B, int n, double dt, double softening, int n_steps,
                       int save_interval, double* sp, double* sv, double* sa, int* final_in_a, cudaStream_t st) {
    NB_REQUIRE(sa_ && sb_ && vel && acc && B > 0 && n > 0 && n <= kBatchedMaxBodies,
               "nb_run_batched: need pointers, B > 0 and 0 < n <= %d", kBatchedMaxBodies);
    NB_REQUIRE(n_steps >= 0 && save_interval >= 1, "n_steps >= 0 and save_interval >= 1 required");
    NB_REQUIRE((sp == nullptr) == (sv == nullptr) && (sp == nullptr) == (sa == nullptr),
               "snapshot outputs must be all set or all null");
    const size_t stream_stride = (size_t)nb_padded_bodies(n) * 4, row = (size_t)n * 3;
    const size_t n_snap = 1 + (size_t)(n_steps / save_interval), snap_stride = n_snap * row;
    const dim3 grid(ceil_div(n, 256), B);
    if (sp) {
        batched_snapshot0_kernel<T><<<grid, 256, 0, st>>>(sa_, vel, acc, n, stream_stride, snap_stride, sp, sv, sa);
        if (int rc = check_launch("batched snapshot0 kernel")) return rc;
    }
    T* cur = sa_;
    T* next = sb_;
    if (n_steps > 0) {
        batched_kick_drift_kernel<T><<<grid, 256, 0, st>>>(cur, next, vel, acc, n, stream_stride, (T)dt, (T)(0.5 * dt));
        if (int rc = check_launch("batched kick_drift kernel")) return rc;
        T* t = cur; cur = next; next = t;
    }
    const int kp = batched_kp(n);
    size_t snap = 1;
    for (int k = 1; k <= n_steps; ++k) {
        int flags = 0;
        if (k < n_steps) flags |= NB_STEP_CONTINUE;
        const bool save = (k % save_interval) == 0;  // nbody.py:240
        if (save) flags |= NB_STEP_SNAPSHOT;
        const bool w = save && sp;
        if (int rc = group_step<T>(cur, next, vel, acc, n, dt, softening, /*mode=*/2, flags, w ? sp + snap * row : nullptr,
                                   w ? sv + snap * row : nullptr, w ? sa + snap * row : nullptr, nullptr, kp, st, B,
                                   snap_stride))
            return rc;
        if (save) ++snap;
        if (k < n_steps) { T* t = cur; cur = next; next = t; }
    }
    if (final_in_a) *final_in_a = (cur == sa_) ? 1 : 0;
    return NB_OK;
}

}  // namespace nb

extern "C" {

int nb_batched_max_bodies(void) { return nb::kBatchedMaxBodies; }

int nb_accel_batched_f64(const double* streams, int B, int n, double softening, double* acc, nb_stream_t s) {
    return nb::accel_batched<double>(streams, B, n, softening, acc, (cudaStream_t)s);
}
int nb_accel_batched_f32(const float* streams, int B, int n, double softening, float* acc, nb_stream_t s) {
    return nb::accel_batched<float>(streams, B, n, softening, acc, (cudaStream_t)s);
}
int nb_run_batched_f64(double* stream_a, double* stream_b, double* vel, double* acc, int B, int n, double dt,
                       double softening, int n_steps, int save_interval, double* snap_pos, double* snap_vel,
                       double* snap_acc, int* final_in_a, nb_stream_t s) {
    return nb::run_batched<double>(stream_a, stream_b, vel, acc, B, n, dt, softening, n_steps, save_interval, snap_pos,
                                   snap_vel, snap_acc, final_in_a, (cudaStream_t)s);
}
int nb_run_batched_f32(float* stream_a, float* stream_b, float* vel, float* acc, int B, int n, double dt,
                       double softening, int n_steps, int save_interval, double* snap_pos, double* snap_vel,
                       double* snap_acc, int* final_in_a, nb_stream_t s) {
    return nb::run_batched<float>(stream_a, stream_b, vel, acc, B, n, dt, softening, n_steps, save_interval, snap_pos,
                                  snap_vel, snap_acc, final_in_a, (cudaStream_t)s);
}

}  // extern "C"
