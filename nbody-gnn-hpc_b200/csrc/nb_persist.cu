// nb_persist.cu -- K2p: ALL leapfrog steps of a run of one mid-size system (640 < N <= 32,768) in ONE launch (sm_100a).
//
// Replaces the loop of NBodySimulator.run, reference src/hpc/nbody.py:237-241, where the per-step kernels (K2,
// nb_force.cu) pay a fixed ~14 us per step -- launch, programmatic-dependent-launch hand-over, mbarrier / TMA prologue
// of every CTA, partial flush, tile counter, last-CTA read-back -- on steps that are 4 .. 100 us of arithmetic, and a
// grid of (i-tiles x j-segments) CTAs that fills 148 SMs in 1.7 waves at N = 16,384.
//
// Decomposition.  The unit of work is a WARP-TASK: 32 lanes x kP bodies (a "group") against ONE j-segment of K2's
// segment plan (nb_segment_plan: a function of N alone).  A task streams its segment global -> shared memory through
// the warp's OWN two-stage ring of 2 KB tiles (1-D bulk TMA issued by lane 0, completion on the warp's own mbarriers:
// warps never meet at a CTA barrier inside a force pass) and applies every tile to its bodies with the very loops of
// K2 (nb_tiles.cuh), so a body's segment partial has K2's bits.  The groups x segments tasks are laid on one line,
// group-major, and cut into equal contiguous ranges, one per CTA of a persistent cooperative grid (2 CTAs per SM,
// <= 14 warps each); a CTA's warps take its tasks round robin.  At N = 16,384 that is 4096 tasks on 296 CTAs = 13.8
// per CTA: one task per warp, every SM within one task of every other -- no wave quantisation.
//
// The warp that completes a group LAST (arrival counter per group) adds the group's segment partials in ascending
// order and integrates its bodies -- closing kick, snapshot row, next opening kick, drift into the other stream
// buffer -- exactly the arithmetic of K2's epilogue: the trajectories are bit-identical to the per-step kernels'.
// One grid-wide barrier per step (release add + acquire poll on one word) orders the next force pass after every
// body's new position; the streams alternate as in K2.  Everything mutable is read at L2 (__ldcg / TMA): the L1 of an
// SM is not coherent with what other SMs wrote a step ago.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "nb_warptask.cuh"

namespace nb {

// Two shapes.  kP = 4 bodies per lane: ONE CTA per SM, up to 16 warps, up to 128 registers -- the inner loop of K2's big
// tile (half the shared-memory loads per interaction of kP = 2).  kP = 2 / 1: TWO CTAs per SM, up to 14 warps each,
// 72 registers -- more, smaller tasks for the smaller systems.
constexpr int kPersistMaxBodies = 32768;
template <int kP> struct PShape { static constexpr int warps = 14, ctas_per_sm = 2; };
template <> struct PShape<4> { static constexpr int warps = 16, ctas_per_sm = 1; };

template <typename T>
struct PersistArgs {
    T* stream_a;         // x_k of step k = 1 is here on entry (after the opening kick + drift)
    T* stream_b;
    T* vel;              // (n,3): v_{1/2} on entry, v_n on return
    T* acc;              // (n,3): a_n on return
    T* partial;          // [n_seg][3][n]
    int* group_counter;  // one word per group, zero on entry, left zero
    unsigned* barrier;   // one word, zero on entry
    int* error;          // the workspace's error word
    int n, n_pad, seg_len, n_seg, n_groups;
    int n_steps, save_interval;
    T dt, half_dt, eps2;
    double* sp;          // snapshot stacks (n_snap, n, 3) or null; row 0 (the entry state) is written by the caller
    double* sv;
    double* sa;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long persist_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Grid-wide barrier number `index` (1, 2, ...) of the launch: every CTA adds one to the word (release) and polls it
// (acquire) until all have.  Returns false if the others never arrive (a bug or a lost co-residency guarantee): the
// caller records it and leaves instead of hanging the GPU.
__device__ __forceinline__ bool grid_barrier(unsigned* word, unsigned index, unsigned long long timeout_ns) {
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = index * gridDim.x;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(word) : "memory");
        unsigned seen;
        const unsigned long long t0 = persist_ns();
        int ok = 1;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(word) : "memory");
            if (seen >= target) break;
            if (persist_ns() - t0 > timeout_ns) { ok = 0; break; }
        } while (true);
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// Body li at step k: its segment partials in ascending order, closing kick, snapshot row, and (cont) the next opening
// kick and the drift into the other stream -- the arithmetic of K2's finish_body, operation for operation
// (reference nbody.py:205-214, NumPy order: every product and sum rounded once).
template <typename T>
__device__ __forceinline__ void persist_finish_body(const PersistArgs<T>& g, const T* __restrict__ cur,
                                                    T* __restrict__ next, int li, bool cont, double* sp, double* sv,
                                                    double* sa) {
    const int n = g.n;
    T a[3] = {T(0), T(0), T(0)};
    for (int s = 0; s < g.n_seg; ++s) {
        const T* p = g.partial + (size_t)s * 3 * n;
#pragma unroll
        for (int c = 0; c < 3; ++c) a[c] += __ldcg(p + (size_t)c * n + li);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T x = __ldcg(cur + StreamIO<T>::index(li, c));
        T v = __ldcg(g.vel + (size_t)li * 3 + c);
        v = mul_add_unfused(g.half_dt, a[c], v);  // closing kick, nbody.py:214
        if (sp) {                                 // get_state(), nbody.py:250-259
            sp[(size_t)li * 3 + c] = (double)x;
            sv[(size_t)li * 3 + c] = (double)v;
            sa[(size_t)li * 3 + c] = (double)a[c];
        }
        if (cont) {
            v = mul_add_unfused(g.half_dt, a[c], v);  // next step's opening kick, nbody.py:205
            x = mul_add_unfused(g.dt, v, x);          // drift, nbody.py:208
            next[StreamIO<T>::index(li, c)] = x;
        }
        g.vel[(size_t)li * 3 + c] = v;
        g.acc[(size_t)li * 3 + c] = a[c];
    }
}

template <typename T, int kP, bool kZeroEps>
__global__ void __launch_bounds__(PShape<kP>::warps * 32, PShape<kP>::ctas_per_sm)
persist_kernel(const PersistArgs<T> g) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    char* ring = smem + (size_t)warp * kPStages * kPTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_warps * kPStages * kPTileBytes) + warp * kPStages;
    if (lane == 0) {
        for (int s = 0; s < kPStages; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    __syncwarp();
    uint32_t tiles_done = 0;  // tiles this warp has consumed since the launch began (ring slot and barrier phase)

    // this CTA's range of the group-major task line
    const long n_tasks = (long)g.n_groups * g.n_seg;
    const int t_lo = (int)(n_tasks * blockIdx.x / gridDim.x), t_hi = (int)(n_tasks * (blockIdx.x + 1) / gridDim.x);
    const size_t row = (size_t)g.n * 3;

    T* cur = g.stream_a;
    T* next = g.stream_b;
    bool ok = true;
    for (int k = 1; k <= g.n_steps; ++k) {
        const bool cont = k < g.n_steps;
        const bool save = g.sp != nullptr && (k % g.save_interval) == 0;  // nbody.py:240
        double* sp = save ? g.sp + (size_t)(k / g.save_interval) * row : nullptr;
        double* sv = save ? g.sv + (size_t)(k / g.save_interval) * row : nullptr;
        double* sa = save ? g.sa + (size_t)(k / g.save_interval) * row : nullptr;
        // the stream was written through the generic proxy (by other SMs, a barrier ago); TMA reads it
        if (lane == 0) asm volatile("fence.proxy.async;" ::: "memory");
        for (int t = t_lo + warp; t < t_hi && ok; t += n_warps) {
            const int grp = t / g.n_seg, seg = t - grp * g.n_seg;
            const int j0 = seg * g.seg_len, j1 = min(j0 + g.seg_len, g.n_pad);
            NB_CHECK(grp >= 0 && grp < g.n_groups && seg >= 0 && seg < g.n_seg && j0 < j1);
            ok = WarpTask<T, kP, kZeroEps>::run(cur, g.n, grp, j0, j1, g.eps2, g.partial + (size_t)seg * 3 * g.n, g.n, 0,
                                                ring, bars, tiles_done, lane);
            if (!ok) break;
            // the group's last task to finish integrates the group
            __threadfence();
            __syncwarp();
            int last = 0;
            if (lane == 0) last = atomicAdd(g.group_counter + grp, 1) == g.n_seg - 1;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                if (lane == 0) g.group_counter[grp] = 0;  // ready for the next step
                __threadfence();
#pragma unroll
                for (int q = 0; q < kP; ++q) {
                    const int li = grp * 32 * kP + lane + 32 * q;
                    if (li < g.n) persist_finish_body<T>(g, cur, next, li, cont, sp, sv, sa);
                }
            }
        }
        if (!ok && lane == 0) atomicCAS(g.error, 0, NB_PERSIST_STALLED);
        if (cont) {
            if (!grid_barrier(g.barrier, (unsigned)k, g.timeout_ns)) {
                if (threadIdx.x == 0) atomicCAS(g.error, 0, NB_PERSIST_STALLED);
                return;
            }
            if (*reinterpret_cast<volatile int*>(g.error) != 0) return;  // somebody stalled: nobody goes on
            T* t = cur; cur = next; next = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
struct PersistPlan {
    int kP, warps, grid;
};

// The largest kP that still gives (nearly) every warp slot of the machine a task; as many warps per CTA as the busiest
// CTA has tasks (capped by the shape).  NB_PERSIST_KP forces kP (experiments).
static PersistPlan persist_plan(int n, int n_seg, int sms) {
    PersistPlan p;
    p.kP = 1;
    for (int kp : {4, 2}) {
        const int ctas = sms * (kp == 4 ? PShape<4>::ctas_per_sm : PShape<2>::ctas_per_sm);
        const long tasks = (long)ceil_div(n, 32 * kp) * n_seg;
        if (tasks >= 12L * ctas) { p.kP = kp; break; }
    }
    if (const char* e = getenv("NB_PERSIST_KP")) {
        const int kp = atoi(e);
        if (kp == 1 || kp == 2 || kp == 4) p.kP = kp;
    }
    const int ctas = sms * (p.kP == 4 ? PShape<4>::ctas_per_sm : PShape<2>::ctas_per_sm);
    const int wmax = p.kP == 4 ? PShape<4>::warps : PShape<2>::warps;
    const long tasks = (long)ceil_div(n, 32 * p.kP) * n_seg;
    p.grid = (int)(tasks < ctas ? tasks : ctas);
    const int per_cta = (int)((tasks + p.grid - 1) / p.grid);
    p.warps = per_cta < wmax ? per_cta : wmax;
    if (p.warps < 1) p.warps = 1;
    return p;
}

template <typename T>
int persist_run(T* stream_a, T* stream_b, T* vel, T* acc, int n, double dt, double softening, int n_steps,
                int save_interval, double* sp, double* sv, double* sa, T* partial, int* group_counter,
                unsigned* barrier, int* error, cudaStream_t st) {
    int dev = 0, sms = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PersistArgs<T> g;
    g.stream_a = stream_a; g.stream_b = stream_b; g.vel = vel; g.acc = acc;
    g.partial = partial; g.group_counter = group_counter; g.barrier = barrier; g.error = error;
    g.n = n; g.n_pad = nb_padded_bodies(n);
    nb_segment_plan(n, &g.seg_len, &g.n_seg);
    const PersistPlan plan = persist_plan(n, g.n_seg, sms);
    g.n_groups = ceil_div(n, 32 * plan.kP);
    g.n_steps = n_steps; g.save_interval = save_interval;
    g.dt = (T)dt; g.half_dt = (T)(0.5 * dt);  // "0.5 * self.dt" is evaluated first, nbody.py:205
    g.eps2 = (T)(softening * softening);
    g.sp = sp; g.sv = sv; g.sa = sa;
    g.timeout_ns = 5000000000ull;
    const bool zero = !(g.eps2 > T(0));
    void (*kern)(const PersistArgs<T>);
    if (plan.kP == 4) kern = zero ? persist_kernel<T, 4, true> : persist_kernel<T, 4, false>;
    else if (plan.kP == 2) kern = zero ? persist_kernel<T, 2, true> : persist_kernel<T, 2, false>;
    else kern = zero ? persist_kernel<T, 1, true> : persist_kernel<T, 1, false>;
    const size_t smem = (size_t)plan.warps * kPStages * (kPTileBytes + sizeof(uint64_t));
    NB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16 * kPStages * (kPTileBytes + 8))));
    int per_sm = 0;
    NB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.warps * 32, smem));
    NB_REQUIRE((long)per_sm * sms >= plan.grid, "persistent step kernel: %d CTAs cannot be resident at once (%d per SM)",
               plan.grid, per_sm);
    NB_CUDA_OK(cudaMemsetAsync(barrier, 0, sizeof(unsigned), st));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(plan.warps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // the grid barrier needs every CTA resident: guaranteed, or refused
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, g));
    return check_launch("persistent step kernel");
}

template int persist_run<float>(float*, float*, float*, float*, int, double, double, int, int, double*, double*, double*,
                                float*, int*, unsigned*, int*, cudaStream_t);
template int persist_run<double>(double*, double*, double*, double*, int, double, double, int, int, double*, double*,
                                 double*, double*, int*, unsigned*, int*, cudaStream_t);

}  // namespace nb

extern "C" int nb_persist_max_bodies(void) { return nb::kPersistMaxBodies; }
