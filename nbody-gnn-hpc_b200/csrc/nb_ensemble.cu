// nb_ensemble.cu -- K3: B independent small systems advanced inside one launch (sm_100a).
//
// Replaces the process pool of the data-generation script, reference
// scripts/generate_data.py:32-58,142-149: there each worker builds one NBodySimulator and calls
// run(n_steps) (src/hpc/nbody.py:220-248), i.e. n_steps x [step() + get_state()].
//
// A system being advanced lives in shared memory:
//   pos   N x {x,y,z,G*m}      read by every thread in the force phase (broadcast LDS.128)
//   vel, acc   3N each          the integrator's state, (body, component) order = API order
//   part  parts x 3N            force partials, one slab per j-part
// A step of one system is two phases:
//   F  every thread owns TWO bodies (r, r + rows) and one of `parts` contiguous j ranges: 2
//      independent interaction chains per loaded j.  N = 200: 100 rows x 5 parts = 500 threads, 16
//      warps, four per scheduler, so the four FP64 pipes of the SM carry equal load.
//   I  the 3N (body, component) scalars are spread over all threads (a thread keeps the same scalars
//      for the whole launch; their offsets are computed once): partials added in ascending part
//      order, closing kick, snapshot row stored straight to HBM -- consecutive threads store
//      consecutive 8-byte words of the reference's (T+1, N, 3) layout --, next opening kick, drift.
// F needs every thread's I of the previous step and I needs every thread's F: two CTA-wide
// dependencies per step, and phase I is a latency-bound sliver (25 dependent FP64 operations) during
// which the FP64 pipe idles.
//
// Two lanes and integrator warps.  ONE CTA per SM advances TWO systems, "lanes".  Its 16 force
// warps run  F(0) F(1) F(0) F(1) ...  without ever meeting at a CTA barrier; four more warps do
// nothing but phase I: I(0) while the force warps are in F(1), I(1) while they are in F(0).  The
// dependencies are mbarriers with split arrive / wait: a force warp arrives on "F(L) done" (count:
// force warps) and moves on; the integrators wait for it, integrate lane L, and arrive on "I(L)
// done" (count: integrator warps), which the force warps only look at a whole force phase later.  The FP64 pipe never drains at a barrier and phase I is off the critical path: a step
// costs its force phase.  With fewer systems than 2 x SMs there is one lane and every thread takes
// part in both phases (same barriers, counts = all warps).
// (Two co-resident 512-thread CTAs were the earlier way to overlap the phases.  Measured with
// tools/exp_ens.cu: the SM favours the CTA that arrived first -- it finishes 400 steps in 3.1 ms,
// its neighbour in 5.3 ms -- so CTAs advance at unpredictable rates and any work sharing between
// them stalls; and with both lanes integrated by all warps in lock step nothing overlaps either:
// 5.42 ms against 5.57 ms for 296 systems; profiles/r01_exp_ens_*.log.)
//
// Few systems (C x B <= SMs for C = 8, 4 or 2; a single simulation is the reference's README
// default): one system per thread-block CLUSTER of C CTAs (cluster_ensemble_kernel).  Every CTA keeps
// ALL positions of its system in its own shared memory (two buffers), owns 1/C of the bodies -- force phase over
// all j from local shared memory, integrate phase for its own bodies -- and stores the drifted
// positions of its bodies straight into the next buffer of all C CTAs through distributed shared
// memory (st.async, counted on the destination's mbarrier: no cluster-wide barrier per step); nothing
// but snapshots touches global memory.  Same j-parts, same summation order: bit-identical to the
// one-CTA kernel at 40 % of its time per step.
//
// Scheduling.  The grid is persistent, one CTA per SM, `lanes` workers per CTA.  The B x n_steps
// system-steps of the launch are laid on one line, system-major, and cut into equal intervals, one
// per worker.  A system that straddles a cut is shared by two neighbouring workers: worker w runs
// its first steps ("head") BEFORE anything else and parks the state in the in/out arrays behind a
// per-system flag; worker w+1 runs the remaining steps ("tail") AFTER everything else of its own.
// An interval is at least one whole system long, so the head has been done for a whole system's
// time when the tail is wanted: all workers advance in lock step (same code, one CTA per SM), so
// nobody waits, every worker gets the same work to within one step, and 300 systems on 296 workers
// cost 300/296 of 296 systems (measured: 5.72 -> 5.80 ms) instead of a second round.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "nb_common.cuh"

namespace nb {

namespace cg = cooperative_groups;

// The mass slot of a body (pos.w): G m -- or, in the first-order float64 build (nb_common.cuh), c0 = 5/2 G m with
// the companion constant c1 = -0.6 c0 kept in the lane's c1[] array.
__device__ __forceinline__ double mass_slot(double gm, double*) { return kMassSlotF64 * gm; }
__device__ __forceinline__ float mass_slot(double gm, float*) { return (float)gm; }

template <bool kZeroEps>
__device__ __forceinline__ void pair_any(double xi, double yi, double zi, double xj, double yj, double zj, double gmj,
                                         double c1j, double eps2, double& ax, double& ay, double& az) {
    pair_f64<kZeroEps>(xi, yi, zi, xj, yj, zj, gmj, c1j, eps2, ax, ay, az);
}
template <bool kZeroEps>
__device__ __forceinline__ void pair_any(float xi, float yi, float zi, float xj, float yj, float zj, float gmj,
                                         float /*c1j*/, float eps2, float& ax, float& ay, float& az) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    float r2 = fmaf(dx, dx, eps2);
    r2 = fmaf(dy, dy, r2);
    r2 = fmaf(dz, dz, r2);
    float inv = rsqrt_approx(r2);
    // i == j (r2 == eps2) contributes exactly 0 whatever the magnitudes: G*m*inv^3 may overflow (see nb_force.cu)
    inv = (r2 > eps2) ? inv : 0.f;
    const float f = (gmj * inv) * (inv * inv);
    ax = fmaf(f, dx, ax);
    ay = fmaf(f, dy, ay);
    az = fmaf(f, dz, az);
}

template <typename T>
struct Vec4;
template <>
struct Vec4<double> {
    using type = double4;
};
template <>
struct Vec4<float> {
    using type = float4;
};

struct EnsembleArgs {
    double* x;  // (B,N,3) in/out
    double* v;
    double* a;
    const void* masses;
    int masses_are_f32;
    int mass_stride;
    int B, N;
    double dt, half_dt, eps2;
    int n_steps, save_interval, compute_a0, write_initial;
    double* out_x;  // (B, n_snap_total, N, 3) or null
    double* out_v;
    double* out_a;
    int n_snap_total, snap_offset;
    int rows;         // threads per part; a thread owns bodies r and r + rows
    int parts;        // j-parts
    int lanes;        // systems advanced side by side by one CTA (1 or 2)
    int f_threads;    // threads that run phase F; blockDim.x - f_threads (0 or 128) threads only integrate
    int* progress;    // per-system flag, zero on entry: 1 once the head steps of a shared system are parked
};

#ifdef NB_F64_PAIR_FIRST_ORDER
constexpr int kC1Slots = 1;  // one more array of N per lane: c1[j]
#else
constexpr int kC1Slots = 0;
#endif

template <typename T>
struct SystemSmem {
    typename Vec4<T>::type* pos;
    T* c1;  // first-order float64 build only
    T* vel;
    T* acc;
    T* part;
};

// shared memory of one lane, rounded to 16 bytes
template <typename T>
__host__ __device__ inline size_t lane_smem_bytes(int N, int parts) {
    const size_t b = (size_t)N * sizeof(typename Vec4<T>::type) + (size_t)((2 + parts) * 3 + kC1Slots) * N * sizeof(T);
    return (b + 15) / 16 * 16;
}

// Thread shape for N bodies: two bodies per thread, as many j-parts as fit in kForceThreadsMax threads (at most 8).
#ifndef NB_K3_FORCE_THREADS
#define NB_K3_FORCE_THREADS 512
#endif
constexpr int kForceThreadsMax = NB_K3_FORCE_THREADS;  // rows x parts never exceeds it (shape_parts)
__host__ __device__ constexpr int shape_rows(int N) { return (N + 1) / 2; }
__host__ __device__ constexpr int shape_parts(int N) {
    int p = kForceThreadsMax / shape_rows(N);
    p = p < 1 ? 1 : p;
    p = p > 8 ? 8 : p;
    return p > N ? N : p;
}

// Which integrator scalars (body, component) a thread owns: idx = first + e * stride over the integrating threads.
// The first two are kept in registers for the whole launch; any further ones are recomputed on the fly.  The same
// thread loads, integrates and stores a scalar, so none of that needs a barrier.
struct Owned {
    int idx[2];   // API-order index 3*i + c, or -1
    int poff[2];  // offset of the coordinate inside pos (4*i + c)
    int first, stride;
};
__device__ __forceinline__ Owned make_owned(int n3, int first, int stride) {
    Owned o;
    o.first = first; o.stride = stride;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int idx = first + e * stride;
        const int i = idx / 3;
        o.idx[e] = idx < n3 ? idx : -1;
        o.poff[e] = idx < n3 ? 4 * i + (idx - 3 * i) : 0;
    }
    return o;
}
template <class F>
__device__ __forceinline__ void for_each_owned(const Owned& own, int n3, F&& f) {
    if (own.idx[0] >= 0) f(own.idx[0], own.poff[0]);
    if (own.idx[1] >= 0) f(own.idx[1], own.poff[1]);
#pragma unroll 4
    for (int idx = own.first + 2 * own.stride; idx < n3; idx += own.stride) {
        const int i = idx / 3;
        f(idx, 4 * i + (idx - 3 * i));
    }
}

// One piece of work of a worker: system b from the synchronised state of step k0 to that of step k1.
struct Piece {
    int b, k0, k1;
    bool wait;     // the state of step k0 was parked by the previous worker: wait for its flag
    bool publish;  // park the state of step k1 for the next worker and raise the flag
};

// A worker's interval [lo, hi) of the B x n system-step line, as the ordered list: head of the system shared with
// the next worker, whole systems, tail of the system shared with the previous worker (see the header).
struct Worker {
    int b_first, s_first, b_last, s_last, b_whole, n_steps, stage;
    __host__ __device__ void init(long w, long n_workers, int B, int n_steps_) {
        n_steps = n_steps_;
        const long n = n_steps_ > 0 ? n_steps_ : 1;  // n_steps == 0 (a_0 / first snapshot only): one unit per system
        const long W = (long)B * n;
        const long lo = w * W / n_workers, hi = (w + 1) * W / n_workers;
        b_first = (int)(lo / n); s_first = (int)(lo - (long)b_first * n);
        b_last = (int)(hi / n);  s_last = (int)(hi - (long)b_last * n);
        b_whole = b_first + (s_first > 0 ? 1 : 0);
        stage = 0;
    }
    __host__ __device__ bool next(Piece& p) {
        if (stage == 0) {
            stage = 1;
            if (s_last > 0) { p = Piece{b_last, 0, s_last, false, true}; return true; }
        }
        if (stage == 1) {
            if (b_whole < b_last) { p = Piece{b_whole++, 0, n_steps, false, false}; return true; }
            stage = 2;
        }
        if (stage == 2) {
            stage = 3;
            if (s_first > 0) { p = Piece{b_first, s_first, n_steps, true, false}; return true; }
        }
        return false;
    }
};

// mbarrier wait that traps instead of hanging the GPU if the other warps never arrive (a bug, not a load effect)
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    unsigned spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
// one arrival per warp, after the whole warp's shared-memory accesses
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Global state of system p.b -> the lane's shared memory, by the owner threads (no barrier needed).
template <typename T>
__device__ __forceinline__ void load_piece(const EnsembleArgs& g, const Piece& p, const SystemSmem<T>& s,
                                           const Owned& own, int N) {
    const int n3 = 3 * N;
    const size_t sbase = (size_t)p.b * n3;
    T* pos_s = reinterpret_cast<T*>(s.pos);
    if (p.wait) {  // every thread polls for itself: its own loads below are ordered after the flag
        volatile int* flag = g.progress + p.b;
        unsigned spins = 0;
        while (*flag == 0) {
            __nanosleep(64);
            if (++spins > (1u << 26)) __trap();  // the previous worker died
        }
        __threadfence();
    }
    for_each_owned(own, n3, [&](int idx, int poff) {
        pos_s[poff] = (T)__ldcg(&g.x[sbase + idx]);  // L2 loads: another SM may have parked this state
        s.vel[idx] = (T)__ldcg(&g.v[sbase + idx]);
        s.acc[idx] = (T)__ldcg(&g.a[sbase + idx]);
    });
    for (int i = own.first; i < N; i += own.stride) {
        const size_t mi = (size_t)p.b * g.mass_stride + i;
        const double m = g.masses_are_f32 ? (double)static_cast<const float*>(g.masses)[mi]
                                          : static_cast<const double*>(g.masses)[mi];
        const T slot = mass_slot(kG * m, (T*)nullptr);  // G * masses[j], nbody.py:57
        pos_s[4 * i + 3] = slot;
        if (kC1Slots) s.c1[i] = slot * (T)kC1OverC0;
    }
}

// The lane's shared memory -> global state of system b, by the owner threads.
template <typename T>
__device__ __forceinline__ void store_piece(const EnsembleArgs& g, int b, const SystemSmem<T>& s, const Owned& own,
                                            int N) {
    const int n3 = 3 * N;
    const size_t sbase = (size_t)b * n3;
    const T* pos_s = reinterpret_cast<const T*>(s.pos);
    for_each_owned(own, n3, [&](int idx, int poff) {
        g.x[sbase + idx] = (double)pos_s[poff];
        g.v[sbase + idx] = (double)s.vel[idx];
        g.a[sbase + idx] = (double)s.acc[idx];
    });
    __threadfence();  // visible device-wide before the flag (raised after the lane's next barrier) can be seen
}

// Phase F of one lane: this thread's two bodies against its j-part, partial sums into the lane's slab q.
// kN != 0: the shape is a compile-time constant (kN % shape_parts(kN) == 0), so the j loop has a known trip count.
template <typename T, bool kZeroEps, int kN>
__device__ __forceinline__ void force_phase(const EnsembleArgs& g, const SystemSmem<T>& s) {
    using V4 = typename Vec4<T>::type;
    const int N = kN ? kN : g.N, n3 = 3 * N;
    const int rows = kN ? shape_rows(kN ? kN : 1) : g.rows;
    const int parts = kN ? shape_parts(kN ? kN : 1) : g.parts;
    const int tid = threadIdx.x;
    const int q = tid / rows;      // j-part of this thread
    const int r = tid - q * rows;  // row: bodies r and r + rows
    if (q >= parts) return;
    const int i0 = r, i1 = r + rows;
    const bool has1 = i1 < N;
    const int jb = (int)(((long)q * N) / parts);
    const int je = kN ? jb + kN / shape_parts(kN ? kN : 1) : (int)(((long)(q + 1) * N) / parts);
    const T eps2 = (T)g.eps2;
    const V4 me0 = s.pos[i0];
    const V4 me1 = s.pos[has1 ? i1 : i0];
    T ax0 = 0, ay0 = 0, az0 = 0, ax1 = 0, ay1 = 0, az1 = 0;
#pragma unroll 8
    for (int j = jb; j < je; ++j) {
        const V4 pj = s.pos[j];
        const T c1 = kC1Slots ? s.c1[j] : T(0);
        pair_any<kZeroEps>(me0.x, me0.y, me0.z, pj.x, pj.y, pj.z, pj.w, c1, eps2, ax0, ay0, az0);
        pair_any<kZeroEps>(me1.x, me1.y, me1.z, pj.x, pj.y, pj.z, pj.w, c1, eps2, ax1, ay1, az1);
    }
    NB_CHECK(q >= 0 && q < parts && jb >= 0 && jb <= je && je <= N && i0 < N && (!has1 || i1 < N));
    T* pa = s.part + (size_t)q * n3;
    pa[3 * i0 + 0] = ax0; pa[3 * i0 + 1] = ay0; pa[3 * i0 + 2] = az0;
    if (has1) { pa[3 * i1 + 0] = ax1; pa[3 * i1 + 1] = ay1; pa[3 * i1 + 2] = az1; }
}

// Phase I of one lane at step k of piece p: the state becomes (x_k, v_k, a_k), is snapshotted, and (k < p.k1)
// moves on to the drifted positions of step k + 1.
template <typename T, int kN>
__device__ __forceinline__ void integrate_phase(const EnsembleArgs& g, const Piece& p, int k, const SystemSmem<T>& s,
                                                const Owned& own) {
    const int N = kN ? kN : g.N, n3 = 3 * N;
    const int parts = kN ? shape_parts(kN ? kN : 1) : g.parts;
    const T dt = (T)g.dt, half_dt = (T)g.half_dt;
    T* pos_s = reinterpret_cast<T*>(s.pos);
    const bool entry_state = (k == p.k0) && (p.k0 > 0);  // (x,v,a)_k were finished by whoever ran the head
    const bool do_force = !entry_state && ((k > 0) || g.compute_a0);
    const bool do_close = !entry_state && k > 0;
    const bool do_open = k < p.k1;
    long srow = -1;  // get_state() before the loop and every save_interval steps, nbody.py:235,240-241
    if (g.out_x && !entry_state) {
        if (k == 0) {
            if (g.write_initial) srow = g.snap_offset;
        } else if ((k % g.save_interval) == 0) {
            srow = g.snap_offset + (g.write_initial ? 1 : 0) + (k / g.save_interval - 1);
        }
    }
    const size_t orow = srow >= 0 ? ((size_t)p.b * g.n_snap_total + (size_t)srow) * n3 : 0;
    NB_CHECK(srow < (long)g.n_snap_total && p.b >= 0 && p.b < g.B && k >= p.k0 && k <= p.k1 && p.k1 <= g.n_steps);
    for_each_owned(own, n3, [&](int idx, int poff) {
        NB_CHECK(idx >= 0 && idx < n3 && poff >= 0 && poff < 4 * N);
        T a = s.acc[idx];
        if (do_force) {
            a = s.part[idx];
#pragma unroll
            for (int q = 1; q < 8; ++q)  // parts <= 8; predicated so that every load is in flight before the adds
                if (q < parts) a += s.part[(size_t)q * n3 + idx];
            s.acc[idx] = a;
        }
        T v = s.vel[idx];
        T x = pos_s[poff];
        if (do_close) v = mul_add_unfused(half_dt, a, v);  // closing kick, nbody.py:214
        if (srow >= 0) {                                   // get_state(), nbody.py:250-259: coalesced 8-byte stores
            g.out_x[orow + idx] = (double)x;
            g.out_v[orow + idx] = (double)v;
            g.out_a[orow + idx] = (double)a;
        }
        if (do_open) {
            v = mul_add_unfused(half_dt, a, v);  // opening kick, nbody.py:205
            x = mul_add_unfused(dt, v, x);       // drift, nbody.py:208
            pos_s[poff] = x;
        }
        s.vel[idx] = v;
    });
}

constexpr int kMaxLanes = 2;
static_assert(kForceThreadsMax % 32 == 0 && kForceThreadsMax + 128 <= 1024, "force + integrator warps must fit one CTA");
constexpr int kIntegratorThreads = 128;                // the integrator warps of the two-lane mode
constexpr int kEnsembleThreadsMax = kForceThreadsMax + kIntegratorThreads;

// kSplit: the two-lane build with integrator warps.  Registers are allocated to warps in groups of four, so a 17th
// warp costs as much as a 20th: four integrator warps, 20 warps x 96 registers.  The one-lane build keeps 16 warps
// and up to 128 registers.
template <typename T, bool kZeroEps, int kN, bool kSplit>
__global__ void __launch_bounds__(kSplit ? kEnsembleThreadsMax : kForceThreadsMax, 1)
ensemble_kernel(const EnsembleArgs g) {
    using V4 = typename Vec4<T>::type;
    extern __shared__ __align__(16) char smem[];
    __shared__ __align__(8) uint64_t bar_f[kMaxLanes], bar_i[kMaxLanes];  // "F / I of lane L is done by every warp in it"
    const int N = kN ? kN : g.N, n3 = 3 * N;
    const int parts = kN ? shape_parts(kN ? kN : 1) : g.parts;
    const int lanes = g.lanes;
    // roles: with an integrator warp the first f_threads only run F and the last warp only runs I
    const int i_threads = kSplit ? (int)blockDim.x - g.f_threads : 0;
    const bool split = kSplit;
    const bool does_f = (int)threadIdx.x < g.f_threads;
    const bool does_i = !split || !does_f;
    const Owned own = split ? make_owned(n3, (int)threadIdx.x - g.f_threads, i_threads)
                            : make_owned(n3, (int)threadIdx.x, (int)blockDim.x);

    SystemSmem<T> sm[kMaxLanes];
    Worker wk[kMaxLanes];
    Piece pc[kMaxLanes];
    bool active[kMaxLanes];
    int k[kMaxLanes], publish_b[kMaxLanes];
    unsigned cyc[kMaxLanes];
#pragma unroll
    for (int L = 0; L < kMaxLanes; ++L) {
        char* base = smem + (size_t)L * lane_smem_bytes<T>(N, parts);
        sm[L].pos = reinterpret_cast<V4*>(base);
        sm[L].c1 = reinterpret_cast<T*>(base + (size_t)N * sizeof(V4));
        sm[L].vel = sm[L].c1 + kC1Slots * N;
        sm[L].acc = sm[L].vel + n3;
        sm[L].part = sm[L].acc + n3;
        active[L] = false;
        k[L] = 0; publish_b[L] = -1; cyc[L] = 0;
        if (L < lanes) {
            wk[L].init((long)blockIdx.x * lanes + L, (long)gridDim.x * lanes, g.B, g.n_steps);
            active[L] = wk[L].next(pc[L]);
            if (active[L]) {
                if (does_i) load_piece<T>(g, pc[L], sm[L], own, N);
                k[L] = pc[L].k0;
            }
        }
    }
    if (threadIdx.x == 0) {
        const uint32_t f_warps = g.f_threads >> 5, i_warps = (split ? i_threads : (int)blockDim.x) >> 5;
        for (int L = 0; L < kMaxLanes; ++L) { mbar_init(&bar_f[L], f_warps); mbar_init(&bar_i[L], i_warps); }
    }
    __syncthreads();

    while (active[0] || active[1]) {
        // ---- F slots ---------------------------------------------------------------------------------------------
#pragma unroll
        for (int L = 0; L < kMaxLanes; ++L) {
            if (!active[L] || !does_f) continue;
            if (cyc[L] > 0) {  // this lane's previous I slot is complete (and its state parked, if it had to be)
                mbar_wait_or_trap(&bar_i[L], (cyc[L] - 1) & 1);
                if (publish_b[L] >= 0 && threadIdx.x == 0) atomicExch(g.progress + publish_b[L], 1);
            }
            const bool entry_state = (k[L] == pc[L].k0) && (pc[L].k0 > 0);
            if (!entry_state && (k[L] > 0 || g.compute_a0)) force_phase<T, kZeroEps, kN>(g, sm[L]);
            warp_arrive(&bar_f[L]);
        }
        // ---- I slots (every thread follows the lanes' bookkeeping; only the integrating threads touch data) --------
#pragma unroll
        for (int L = 0; L < kMaxLanes; ++L) {
            if (!active[L]) continue;
            if (cyc[L] > 0) publish_b[L] = -1;  // raised by thread 0 in the F slot above
            if (does_i) {
                mbar_wait_or_trap(&bar_f[L], cyc[L] & 1);  // every force warp has finished this lane's F slot
                integrate_phase<T, kN>(g, pc[L], k[L], sm[L], own);
            }
            if (k[L] < pc[L].k1) {
                ++k[L];
            } else {  // piece finished: park / return the state, take the next piece
                if (does_i) store_piece<T>(g, pc[L].b, sm[L], own, N);
                if (pc[L].publish) publish_b[L] = pc[L].b;
                active[L] = wk[L].next(pc[L]);
                if (active[L]) {
                    if (pc[L].wait && (publish_b[0] >= 0 || publish_b[1] >= 0)) {
                        // The awaited flag may be one this very CTA still has to raise (the other lane's head, in
                        // runs of a few steps): raise what is pending before anybody polls.
                        __syncthreads();
                        if (threadIdx.x == 0)
                            for (int M = 0; M < kMaxLanes; ++M)
                                if (publish_b[M] >= 0) atomicExch(g.progress + publish_b[M], 1);
                        publish_b[0] = publish_b[1] = -1;
                    }
                    if (does_i) load_piece<T>(g, pc[L], sm[L], own, N);
                    k[L] = pc[L].k0;
                }
            }
            if (does_i) warp_arrive(&bar_i[L]);
            ++cyc[L];
        }
    }
    // a head is always followed by other pieces of the same worker, so no flag is pending here; be safe anyway
    __syncthreads();
    if (threadIdx.x == 0)
        for (int L = 0; L < kMaxLanes; ++L)
            if (publish_b[L] >= 0) atomicExch(g.progress + publish_b[L], 1);
}

// ------------------------------------------------------------------------------------------------------------------
// One system per cluster of C = 8, 4 or 2 CTAs (few systems: the largest C with C x B <= SMs).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kClusterCtas = 8;  // the portable maximum; the launch may use 4 or 2

// shared memory of one CTA of a cluster: two full position buffers, partial slabs / vel / acc of its own slab
template <typename T>
__host__ __device__ inline size_t cluster_smem_bytes(int N, int parts, int C) {
    const int S = (N + C - 1) / C;
    return 2 * (size_t)N * sizeof(typename Vec4<T>::type) + (size_t)kC1Slots * N * sizeof(T) +
           (size_t)(parts + 2) * 3 * S * sizeof(T);
}

// --- distributed shared memory with transaction counting ----------------------------------------------------------
// A drifted coordinate goes to every CTA of the cluster as st.async: a remote shared-memory store whose bytes are
// counted on the DESTINATION CTA's mbarrier.  The destination waits for "3N coordinates have arrived" instead of for a
// cluster-wide barrier -- barrier.cluster.arrive.release would also wait for this step's snapshot stores to global
// memory to drain (measured: 3.9 us per step with snapshots against 2.5 us without).
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, int cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_async(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void st_async(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_bar)
                 : "memory");
}

// kMaxThreads bounds the launch so that small slabs (N = 200: 128 threads) get the registers to keep many
// interactions in flight: with one warp per scheduler the force loop is a latency chain, not a throughput problem.
template <typename T, bool kZeroEps, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads, 1) cluster_ensemble_kernel(const EnsembleArgs g) {
    using V4 = typename Vec4<T>::type;
    extern __shared__ __align__(16) char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int C = (int)cluster.num_blocks();                    // cudaLaunchAttributeClusterDimension of the launch
    const int n_clusters = gridDim.x / C, cid = blockIdx.x / C;
    const int N = g.N, n3 = 3 * N, parts = g.parts;
    const int S = (N + C - 1) / C;                              // bodies per CTA
    const int i_lo = min(N, rank * S), ns = min(N, i_lo + S) - i_lo;
    const int tid = threadIdx.x;
    V4* const pos0 = reinterpret_cast<V4*>(smem);              // buffer c is pos0 + c * N
    T* c1s = reinterpret_cast<T*>(pos0 + 2 * N);                // N, first-order float64 build only
    T* part = c1s + kC1Slots * N;                               // parts x 3S
    T* vel = part + (size_t)parts * 3 * S;                      // 3S, (body, component) order
    T* acc = vel + 3 * S;
    // every CTA's copy of the two position buffers, as seen from here (distributed shared memory)
    // every CTA's position buffers and arrival barriers as seen from here (shared::cluster addresses)
    __shared__ __align__(8) uint64_t arrived[2];   // arrived[c]: all 3N coordinates of buffer c are in
    uint32_t remote_pos[kClusterCtas], remote_bar[kClusterCtas];
#pragma unroll
    for (int t = 0; t < kClusterCtas; ++t) {
        remote_pos[t] = map_to_cta(smem_u32(pos0), t < C ? t : 0);       // buffer 0; buffer 1 follows at + N * sizeof(V4)
        remote_bar[t] = map_to_cta(smem_u32(&arrived[0]), t < C ? t : 0);  // arrived[1] follows at + 8
    }
    if (tid == 0) {
        mbar_init(&arrived[0], 1);
        mbar_init(&arrived[1], 1);
        mbar_init_fence();
    }
    uint32_t phase0 = 0u, phase1 = 0u;  // completed uses of arrived[0] / arrived[1], for the wait parity
    const T dt = (T)g.dt, half_dt = (T)g.half_dt, eps2 = (T)g.eps2;
    // force phase: one body of the slab and one j-part per thread (the j-parts of the one-CTA kernel: same sums)
    const int q = tid / S, li_f = tid - q * S;
    const bool f_active = q < parts && li_f < ns;
    const int jb = (int)(((long)q * N) / parts), je = (int)(((long)(q + 1) * N) / parts);

    for (int b = cid; b < g.B; b += n_clusters) {
        const size_t sbase = (size_t)b * n3;
        // all positions and G*m into both local buffers; velocity and acceleration of the own slab
        for (int idx = tid; idx < n3; idx += blockDim.x) {
            const int i = idx / 3, c = idx - 3 * i;
            const T xv = (T)__ldcg(&g.x[sbase + idx]);
            reinterpret_cast<T*>(pos0)[4 * i + c] = xv;
            reinterpret_cast<T*>(pos0 + N)[4 * i + c] = xv;
        }
        for (int i = tid; i < N; i += blockDim.x) {
            const size_t mi = (size_t)b * g.mass_stride + i;
            const double m = g.masses_are_f32 ? (double)static_cast<const float*>(g.masses)[mi]
                                              : static_cast<const double*>(g.masses)[mi];
            const T gm = mass_slot(kG * m, (T*)nullptr);  // G * masses[j], nbody.py:57
            reinterpret_cast<T*>(pos0)[4 * i + 3] = gm;
            reinterpret_cast<T*>(pos0 + N)[4 * i + 3] = gm;
            if (kC1Slots) c1s[i] = gm * (T)kC1OverC0;
        }
        for (int idx = tid; idx < 3 * ns; idx += blockDim.x) {
            vel[idx] = (T)__ldcg(&g.v[sbase + 3 * i_lo + idx]);
            acc[idx] = (T)__ldcg(&g.a[sbase + 3 * i_lo + idx]);
        }
        // nobody stores into a neighbour's buffer before that neighbour has finished loading (and finished the
        // previous system, and initialised its barriers)
        cluster.sync();
        int cur = 0;
        for (int k = 0; k <= g.n_steps; ++k) {
            const bool do_force = (k > 0) || g.compute_a0;
            const bool do_close = k > 0;
            const bool do_open = k < g.n_steps;
            long srow = -1;  // get_state() before the loop and every save_interval steps, nbody.py:235,240-241
            if (g.out_x) {
                if (k == 0) {
                    if (g.write_initial) srow = g.snap_offset;
                } else if ((k % g.save_interval) == 0) {
                    srow = g.snap_offset + (g.write_initial ? 1 : 0) + (k / g.save_interval - 1);
                }
            }
            if (do_open && tid == 0)  // this step's 3N drifted coordinates will land in the other buffer
                mbar_arrive_expect_tx(&arrived[cur ^ 1], (uint32_t)(n3 * sizeof(T)));
            if (do_force) {
                if (f_active) {
                    const V4* pos = pos0 + (size_t)cur * N;
                    const V4 me = pos[i_lo + li_f];
                    T ax = 0, ay = 0, az = 0;
                    // one warp per scheduler and one body per thread: the only parallelism is across j, so eight
                    // interactions are kept in flight (the sums still run in ascending j)
#pragma unroll 20
                    for (int j = jb; j < je; ++j) {
                        const V4 pj = pos[j];
                        pair_any<kZeroEps>(me.x, me.y, me.z, pj.x, pj.y, pj.z, pj.w, kC1Slots ? c1s[j] : T(0), eps2, ax,
                                           ay, az);
                    }
                    T* pa = part + (size_t)q * 3 * S + 3 * li_f;
                    pa[0] = ax; pa[1] = ay; pa[2] = az;
                }
                __syncthreads();
            }
            const size_t orow = srow >= 0 ? ((size_t)b * g.n_snap_total + (size_t)srow) * n3 + 3 * i_lo : 0;
            const T* pos_cur = reinterpret_cast<const T*>(pos0 + (size_t)cur * N);
            for (int idx = tid; idx < 3 * ns; idx += blockDim.x) {
                const int li = idx / 3, c = idx - 3 * li;
                T a = acc[idx];
                if (do_force) {
                    a = part[idx];
                    for (int p = 1; p < parts; ++p) a += part[(size_t)p * 3 * S + idx];
                    acc[idx] = a;
                }
                T v = vel[idx];
                T x = pos_cur[4 * (i_lo + li) + c];
                if (do_close) v = mul_add_unfused(half_dt, a, v);  // closing kick, nbody.py:214
                if (srow >= 0) {
                    g.out_x[orow + idx] = (double)x;
                    g.out_v[orow + idx] = (double)v;
                    g.out_a[orow + idx] = (double)a;
                }
                if (do_open) {
                    v = mul_add_unfused(half_dt, a, v);  // opening kick, nbody.py:205
                    x = mul_add_unfused(dt, v, x);       // drift, nbody.py:208
                    NB_CHECK(i_lo + li < N && c < 3 && (cur == 0 || cur == 1));
                    const uint32_t off = (uint32_t)(((size_t)(cur ^ 1) * 4 * N + 4 * (i_lo + li) + c) * sizeof(T));
#pragma unroll
                    for (int t = 0; t < kClusterCtas; ++t)
                        if (t < C) st_async(remote_pos[t] + off, x, remote_bar[t] + 8u * (uint32_t)(cur ^ 1));
                }
                vel[idx] = v;
            }
            if (do_open) {
                // The next buffer is complete HERE once its 3N coordinates have arrived.  Nobody can overwrite the
                // buffer just read before everybody is done with it: writing it again takes a full force phase on
                // the other buffer, which needs every CTA's coordinates of this step.
                cur ^= 1;
                mbar_wait_or_trap(arrived + cur, (cur ? phase1 : phase0) & 1u);
                if (cur) ++phase1; else ++phase0;
            }
        }
        // final state of the own slab
        const T* pos_fin = reinterpret_cast<const T*>(pos0 + (size_t)cur * N);
        for (int idx = tid; idx < 3 * ns; idx += blockDim.x) {
            const int li = idx / 3, c = idx - 3 * li;
            g.x[sbase + 3 * i_lo + idx] = (double)pos_fin[4 * (i_lo + li) + c];
            g.v[sbase + 3 * i_lo + idx] = (double)vel[idx];
            g.a[sbase + 3 * i_lo + idx] = (double)acc[idx];
        }
        __syncthreads();  // the local buffers are reloaded for the next system
    }
    cluster.sync();  // no CTA exits while a neighbour may still store into its shared memory
}

constexpr int kEnsembleMaxBodies = 1024;

// ------------------------------------------------------------------------------------------------------------------
// launch plans
// ------------------------------------------------------------------------------------------------------------------
struct LaunchPlan {
    void (*kern)(const EnsembleArgs);
    int cluster;  // CTAs per system (8, 4, 2), or 0: the persistent one-CTA-per-SM kernel
    int threads;
    size_t smem;
    int lanes;    // persistent kernel: systems advanced side by side by one CTA
    int grid;     // persistent kernel: CTAs
};
using PlanKey = std::tuple<int, int, int, int, int>;  // device, sizeof(T), B, N, flags
static std::mutex g_plan_mutex;
static std::map<PlanKey, LaunchPlan> g_plans;

template <typename T>
static int make_plan(int dev, int B, int N, int parts, int f_threads, bool zero, bool no_cluster, LaunchPlan* out) {
    int sms = 0, smem_max = 0;
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // Few systems: one system per cluster of 8, 4 or 2 CTAs (distributed shared memory) -- the largest cluster that
    // still gives every system its own -- if the slab shape fits a CTA.
    for (int C = kClusterCtas; C >= 2 && !no_cluster; C /= 2) {
        if ((long)B * C > sms) continue;
        const int S = ceil_div(N, C);
        const int c_threads = round_up(S * parts > 3 * S ? S * parts : 3 * S, 32);
        const size_t c_smem = cluster_smem_bytes<T>(N, parts, C);
        if (N < 2 * C || c_threads > 1024 || c_smem + 1024 > (size_t)smem_max) continue;
        void (*ck)(const EnsembleArgs);
        if (c_threads <= 256)
            ck = zero ? cluster_ensemble_kernel<T, true, 256> : cluster_ensemble_kernel<T, false, 256>;
        else
            ck = zero ? cluster_ensemble_kernel<T, true, 1024> : cluster_ensemble_kernel<T, false, 1024>;
        // the permission is set to the device maximum once: plans of other N share the kernel
        NB_CUDA_OK(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(B * C);
        cfg.blockDim = dim3(c_threads);
        cfg.dynamicSmemBytes = c_smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        // every system's cluster must be resident at once (clusters are placed inside a GPC, so fewer fit than
        // SMs / C: 37 clusters of 4 ran in two waves, 2.1 ms against 1.8 ms on clusters of 2)
        if (cudaOccupancyMaxActiveClusters(&n_clusters, ck, &cfg) == cudaSuccess && n_clusters >= B) {
            *out = LaunchPlan{ck, C, c_threads, c_smem, 1, 0};
            return NB_OK;
        }
        (void)cudaGetLastError();  // not all clusters of this size can be placed at once: try smaller, then one CTA
    }
    // Two lanes + the integrator warps when there are systems for 2 x SMs workers and both lanes fit in shared
    // memory; else one lane, every thread in both phases.
    const size_t lane_bytes = lane_smem_bytes<T>(N, parts);
    const int lanes = (B >= 2 * sms && 2 * lane_bytes + 1024 <= (size_t)smem_max) ? 2 : 1;
    const int threads = f_threads + (lanes == 2 ? kIntegratorThreads : 0);
    const size_t smem = (size_t)lanes * lane_bytes;
    void (*kern)(const EnsembleArgs);
    const bool split = lanes == 2;
    if (N == 200 && !zero)  // the reference's data-generation shape (generate_data.py:109), compiled with constant bounds
        kern = split ? ensemble_kernel<T, false, 200, true> : ensemble_kernel<T, false, 200, false>;
    else if (split)
        kern = zero ? ensemble_kernel<T, true, 0, true> : ensemble_kernel<T, false, 0, true>;
    else
        kern = zero ? ensemble_kernel<T, true, 0, false> : ensemble_kernel<T, false, 0, false>;
    NB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
    int per_sm = 0;
    NB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    NB_REQUIRE(per_sm >= 1, "ensemble kernel does not fit on an SM (N=%d threads=%d smem=%zu)", N, threads, smem);
    // persistent grid, one CTA per SM; with fewer systems than SMs, one system per CTA
    *out = LaunchPlan{kern, 0, threads, smem, lanes, B < sms ? B : sms};
    return NB_OK;
}

template <typename T>
static int ensemble_impl(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride, int B,
                         int N, double dt, double softening, int n_steps, int save_interval, int compute_a0,
                         int write_initial, double* out_x, double* out_v, double* out_a, int n_snap_total,
                         int snap_offset, void* ws, size_t ws_bytes, cudaStream_t st) {
    NB_REQUIRE(x && v && a && masses, "null pointer argument");
    NB_REQUIRE(B > 0 && N > 0 && N <= kEnsembleMaxBodies, "need B > 0 and 0 < N <= %d (got B=%d N=%d)",
               kEnsembleMaxBodies, B, N);
    NB_REQUIRE(n_steps >= 0 && save_interval >= 1, "n_steps >= 0 and save_interval >= 1 required");
    NB_REQUIRE(mass_stride == 0 || mass_stride == N, "mass_stride must be 0 (shared) or N");
    NB_REQUIRE((out_x == nullptr) == (out_v == nullptr) && (out_x == nullptr) == (out_a == nullptr),
               "snapshot outputs must be all set or all null");
    if (out_x) {
        const int rows = (write_initial ? 1 : 0) + n_steps / save_interval;
        NB_REQUIRE(snap_offset >= 0 && snap_offset + rows <= n_snap_total,
                   "snapshot rows [%d, %d) exceed n_snap_total=%d", snap_offset, snap_offset + rows, n_snap_total);
    }
    EnsembleArgs g;
    g.x = x; g.v = v; g.a = a;
    g.masses = masses; g.masses_are_f32 = masses_are_f32; g.mass_stride = mass_stride;
    g.B = B; g.N = N;
    g.dt = dt; g.half_dt = 0.5 * dt; g.eps2 = softening * softening;
    g.n_steps = n_steps; g.save_interval = save_interval; g.compute_a0 = compute_a0; g.write_initial = write_initial;
    g.out_x = out_x; g.out_v = out_v; g.out_a = out_a;
    g.n_snap_total = n_snap_total; g.snap_offset = snap_offset;
    g.rows = shape_rows(N);
    const int parts = shape_parts(N);
    g.parts = parts;
    static_assert(200 % shape_parts(200) == 0, "the static N=200 shape needs equal j-parts");
    g.f_threads = round_up(g.rows * parts, 32);
    const bool zero = !((T)g.eps2 > T(0));
    int dev = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    // Which kernel, cluster size, block size and shared memory this (device, precision, B, N) gets is decided once
    // (occupancy queries and function attributes cost more host time than a one-step launch) and remembered.
    const bool no_cluster = getenv("NB_ENSEMBLE_NO_CLUSTER") != nullptr;
    const PlanKey key{dev, (int)sizeof(T), B, N, (zero ? 1 : 0) | (no_cluster ? 2 : 0)};
    LaunchPlan plan;
    bool have = false;
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        auto it = g_plans.find(key);
        if (it != g_plans.end()) { plan = it->second; have = true; }
    }
    if (!have) {
        if (int rc = make_plan<T>(dev, B, N, parts, g.f_threads, zero, no_cluster, &plan)) return rc;
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        g_plans[key] = plan;
    }
    g.lanes = plan.lanes;
    g.progress = nullptr;
    if (plan.cluster > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(B * plan.cluster);
        cfg.blockDim = dim3(plan.threads);
        cfg.dynamicSmemBytes = plan.smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = plan.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        NB_CUDA_OK(cudaLaunchKernelEx(&cfg, plan.kern, g));
        return check_launch("cluster ensemble kernel");
    }
    // A system is shared by two workers whenever the B x n_steps line does not divide evenly: per-system flags
    if (B % (plan.grid * g.lanes) != 0) {
        NB_REQUIRE(ws && ws_bytes >= nb_ensemble_workspace_bytes(B), "ensemble workspace too small: %zu < %zu",
                   ws_bytes, nb_ensemble_workspace_bytes(B));
        g.progress = static_cast<int*>(ws);
        NB_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)B, st));
    }
    // Workers hand systems to each other through the flags and spin on them: every CTA of the grid must be resident
    // at once.  A COOPERATIVE launch guarantees that (the grid waits until it can be placed as a whole, and a grid
    // that can never be is refused with an error) -- a plain launch next to another stream's kernels or under MPS
    // would leave tail workers spinning for CTAs that are not running.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(plan.threads);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = g.progress != nullptr ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NB_CUDA_OK(cudaLaunchKernelEx(&cfg, plan.kern, g));
    return check_launch("ensemble kernel");
}

}  // namespace nb

extern "C" {

int nb_ensemble_max_bodies(void) { return nb::kEnsembleMaxBodies; }

// The pieces worker w of n_workers would run, in order (host arithmetic; what the kernel's workers compute for
// themselves): out[5*i .. 5*i+4] = {system, first step, last step, waits for the previous worker, parks for the next}.
int nb_ensemble_worker_plan(int B, int n_steps, int n_workers, int w, int* out, int max_pieces, int* n_pieces) {
    NB_REQUIRE(B > 0 && n_steps >= 0 && n_workers > 0 && n_workers <= B && w >= 0 && w < n_workers && out && n_pieces,
               "need 0 < n_workers <= B, 0 <= w < n_workers, n_steps >= 0 and output pointers");
    nb::Worker wk;
    wk.init(w, n_workers, B, n_steps);
    nb::Piece p;
    int n = 0;
    while (wk.next(p)) {
        if (n < max_pieces) {
            out[5 * n + 0] = p.b; out[5 * n + 1] = p.k0; out[5 * n + 2] = p.k1;
            out[5 * n + 3] = p.wait ? 1 : 0; out[5 * n + 4] = p.publish ? 1 : 0;
        }
        ++n;
    }
    *n_pieces = n;  // may exceed max_pieces: only the first max_pieces were written
    return NB_OK;
}

// per-system hand-over flags
size_t nb_ensemble_workspace_bytes(int B) { return (sizeof(int) * ((size_t)(B > 0 ? B : 0) + 1) + 255) / 256 * 256; }

int nb_ensemble_f64(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride, int B,
                    int N, double dt, double softening, int n_steps, int save_interval, int compute_a0,
                    int write_initial, double* out_x, double* out_v, double* out_a, int n_snap_total, int snap_offset,
                    void* ws, size_t ws_bytes, nb_stream_t s) {
    return nb::ensemble_impl<double>(x, v, a, masses, masses_are_f32, mass_stride, B, N, dt, softening, n_steps,
                                     save_interval, compute_a0, write_initial, out_x, out_v, out_a, n_snap_total,
                                     snap_offset, ws, ws_bytes, (cudaStream_t)s);
}
int nb_ensemble_f32(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride, int B,
                    int N, double dt, double softening, int n_steps, int save_interval, int compute_a0,
                    int write_initial, double* out_x, double* out_v, double* out_a, int n_snap_total, int snap_offset,
                    void* ws, size_t ws_bytes, nb_stream_t s) {
    return nb::ensemble_impl<float>(x, v, a, masses, masses_are_f32, mass_stride, B, N, dt, softening, n_steps,
                                    save_interval, compute_a0, write_initial, out_x, out_v, out_a, n_snap_total,
                                    snap_offset, ws, ws_bytes, (cudaStream_t)s);
}

}  // extern "C"
