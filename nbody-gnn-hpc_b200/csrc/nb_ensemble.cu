// nb_ensemble.cu -- K3: B independent small systems advanced inside one launch (sm_100a).
//
// Replaces the process pool of the data-generation script, reference
// scripts/generate_data.py:32-58,142-149: there each worker builds one NBodySimulator and calls
// run(n_steps) (src/hpc/nbody.py:220-248), i.e. n_steps x [step() + get_state()].
//
// One CTA owns one system for a range of steps; the whole state lives in shared memory:
//   pos   N x {x,y,z,G*m}      read by every thread in the force phase (broadcast LDS.128)
//   vel, acc   3N each          the integrator's state, (body, component) order = API order
//   part  parts x 3N            force partials, one slab per j-part
//   stage 3 x 3N                the snapshot row (x_k, v_k, a_k) waiting to be streamed out
// A step is two phases separated by two CTA barriers:
//   F  every thread owns TWO bodies (r, r + rows) and one of `parts` contiguous j ranges: 2
//      independent interaction chains per loaded j.  Before its force loop the thread streams its
//      share of the previous step's snapshot row from `stage` to HBM -- consecutive 8-byte words,
//      fully coalesced, in the reference's (T+1, N, 3) layout, overlapped with the arithmetic.
//   I  the 3N (body, component) scalars are spread over all threads: partials added in ascending
//      part order, closing kick, snapshot into `stage`, next opening kick and drift.
// For N = 200 this is 100 rows x 5 parts = 500 threads -> 16 warps, four per scheduler, two CTAs
// per SM: the four FP64 pipes of an SM carry equal load (13-warp CTAs lost 20% to that skew).
//
// Scheduling.  The grid is persistent: at most (resident CTAs per SM) x (SM count) CTAs.  Whole
// grid-rounds of systems are "home" systems: a CTA keeps its system in shared memory from the first
// step to the last.  The B mod grid leftover systems (4 of 300 on 148 SMs) advance in step chunks
// that home CTAs steal at their own (staggered) step boundaries -- (chunk, system) tickets from a
// global counter, chunk-major; a system's state is handed from chunk to chunk through the in/out
// state arrays and a per-system progress word.  With many leftovers every system advances by
// tickets.  A ticket's predecessor always has a lower ticket number, hence a CTA that is already
// running, so the waits cannot deadlock.  This removes the 300-systems-on-148-SMs tail.
#include "nb_common.cuh"

namespace nb {

template <bool kZeroEps>
__device__ __forceinline__ void pair_any(double xi, double yi, double zi, double xj, double yj, double zj, double gmj,
                                         double eps2, double& ax, double& ay, double& az) {
    pair_f64<kZeroEps>(xi, yi, zi, xj, yj, zj, gmj, eps2, ax, ay, az);
}
template <bool kZeroEps>
__device__ __forceinline__ void pair_any(float xi, float yi, float zi, float xj, float yj, float zj, float gmj,
                                         float eps2, float& ax, float& ay, float& az) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    float r2 = fmaf(dx, dx, eps2);
    r2 = fmaf(dy, dy, r2);
    r2 = fmaf(dz, dz, r2);
    float inv = rsqrt_approx(r2);
    if (kZeroEps) inv = (r2 > 0.f) ? inv : 0.f;
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
    int chunk_steps;  // steps per leftover ticket
    int n_chunks;
    int* sm_slots;    // per-SM arrival counter (zero on entry) used to stagger co-resident CTAs, or null
    unsigned stagger_ns;  // delay of every second CTA of an SM, about half a step
    int home_rounds;  // whole grid-rounds of systems that stay resident in their CTA (0: everything by ticket)
    int* ticket;      // ticket mode: global ticket counter, zero on entry
    int* progress;    // ticket mode: per-system count of finished chunks, zero on entry
};

template <typename T>
struct SystemSmem {
    typename Vec4<T>::type* pos;
    T* vel;
    T* acc;
    T* part;
    T* stage;  // x | v | a, 3N each
};

template <typename T>
__host__ __device__ inline size_t ensemble_smem_bytes(int N, int parts) {
    return (size_t)N * sizeof(typename Vec4<T>::type) + (size_t)(2 + parts + 3) * 3 * N * sizeof(T);
}

// Stream the staged snapshot row to HBM: 3 x 3N consecutive doubles, coalesced.
template <typename T>
__device__ __forceinline__ void flush_stage(const EnsembleArgs& g, const SystemSmem<T>& s, int b, long srow) {
    const int n3 = 3 * g.N;
    const size_t o = ((size_t)b * g.n_snap_total + (size_t)srow) * n3;
    for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
        g.out_x[o + idx] = (double)s.stage[idx];
        g.out_v[o + idx] = (double)s.stage[n3 + idx];
        g.out_a[o + idx] = (double)s.stage[2 * n3 + idx];
    }
}

// Global state of system b -> shared memory (synchronised state of some step k).
template <typename T>
__device__ __forceinline__ void load_state(const EnsembleArgs& g, int b, const SystemSmem<T>& s) {
    const int N = g.N, n3 = 3 * N;
    const int tid = threadIdx.x;
    const size_t sbase = (size_t)b * n3;
    T* pos_s = reinterpret_cast<T*>(s.pos);
    __syncthreads();  // the previous system's readers are done with the shared state
    for (int idx = tid; idx < n3; idx += blockDim.x) {
        // L2 loads: another SM may have written this state a chunk ago
        const int i = idx / 3, c = idx - 3 * i;
        pos_s[4 * i + c] = (T)__ldcg(&g.x[sbase + idx]);
        s.vel[idx] = (T)__ldcg(&g.v[sbase + idx]);
        s.acc[idx] = (T)__ldcg(&g.a[sbase + idx]);
    }
    for (int i = tid; i < N; i += blockDim.x) {
        const size_t mi = (size_t)b * g.mass_stride + i;
        const double m = g.masses_are_f32 ? (double)static_cast<const float*>(g.masses)[mi]
                                          : static_cast<const double*>(g.masses)[mi];
        pos_s[4 * i + 3] = (T)(kG * m);  // G * masses[j], nbody.py:57
    }
    __syncthreads();
}

// Shared memory -> global state of system b, published for other CTAs.
template <typename T>
__device__ __forceinline__ void store_state(const EnsembleArgs& g, int b, const SystemSmem<T>& s) {
    const int n3 = 3 * g.N;
    const size_t sbase = (size_t)b * n3;
    const T* pos_s = reinterpret_cast<const T*>(s.pos);
    for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
        const int i = idx / 3, c = idx - 3 * i;
        g.x[sbase + idx] = (double)pos_s[4 * i + c];
        g.v[sbase + idx] = (double)s.vel[idx];
        g.a[sbase + idx] = (double)s.acc[idx];
    }
    __threadfence();  // visible before a progress word is advanced
}

// Advance the system held in shared memory from step k_begin to k_end (synchronised state in, synchronised
// state out).  k_begin == 0 additionally handles the initial acceleration and the initial snapshot.
template <typename T, bool kZeroEps>
__device__ __forceinline__ void run_steps(const EnsembleArgs& g, int b, int k_begin, int k_end,
                                          const SystemSmem<T>& s) {
    using V4 = typename Vec4<T>::type;
    const int N = g.N, n3 = 3 * N;
    const int tid = threadIdx.x;
    const int q = tid / g.rows;      // j-part of this thread
    const int r = tid - q * g.rows;  // row: bodies r and r + rows
    const bool active = q < g.parts;
    const int i0 = r, i1 = r + g.rows;
    const bool has1 = i1 < N;
    const int jb = active ? (int)(((long)q * N) / g.parts) : 0;
    const int je = active ? (int)(((long)(q + 1) * N) / g.parts) : 0;
    const T dt = (T)g.dt, half_dt = (T)g.half_dt, eps2 = (T)g.eps2;
    T* pos_s = reinterpret_cast<T*>(s.pos);

    long pending = -1;  // snapshot row staged but not yet streamed out (uniform across the CTA)
    for (int k = k_begin; k <= k_end; ++k) {
        const bool entry_state = (k == k_begin) && (k_begin > 0);  // (x,v,a)_k were finished by the previous chunk
        const bool do_force = !entry_state && ((k > 0) || g.compute_a0);
        const bool do_close = !entry_state && k > 0;
        const bool do_open = k < k_end;
        long srow = -1;  // get_state() before the loop and every save_interval steps, nbody.py:235,240-241
        if (g.out_x && !entry_state) {
            if (k == 0) {
                if (g.write_initial) srow = g.snap_offset;
            } else if ((k % g.save_interval) == 0) {
                srow = g.snap_offset + (g.write_initial ? 1 : 0) + (k / g.save_interval - 1);
            }
        }
        // ---- phase F -----------------------------------------------------------------------------
        if (pending >= 0) {
            flush_stage<T>(g, s, b, pending);
            pending = -1;
        }
        if (do_force) {
            if (active) {
                const V4 me0 = s.pos[i0];
                const V4 me1 = s.pos[has1 ? i1 : i0];
                T ax0 = 0, ay0 = 0, az0 = 0, ax1 = 0, ay1 = 0, az1 = 0;
#pragma unroll 4
                for (int j = jb; j < je; ++j) {
                    const V4 pj = s.pos[j];
                    pair_any<kZeroEps>(me0.x, me0.y, me0.z, pj.x, pj.y, pj.z, pj.w, eps2, ax0, ay0, az0);
                    pair_any<kZeroEps>(me1.x, me1.y, me1.z, pj.x, pj.y, pj.z, pj.w, eps2, ax1, ay1, az1);
                }
                T* pa = s.part + (size_t)q * n3;
                pa[3 * i0 + 0] = ax0; pa[3 * i0 + 1] = ay0; pa[3 * i0 + 2] = az0;
                if (has1) { pa[3 * i1 + 0] = ax1; pa[3 * i1 + 1] = ay1; pa[3 * i1 + 2] = az1; }
            }
            __syncthreads();
        }
        // ---- phase I -----------------------------------------------------------------------------
        for (int idx = tid; idx < n3; idx += blockDim.x) {
            const int i = idx / 3, c = idx - 3 * i;
            T a = s.acc[idx];
            if (do_force) {
                a = s.part[idx];
                for (int p = 1; p < g.parts; ++p) a += s.part[(size_t)p * n3 + idx];
                s.acc[idx] = a;
            }
            T v = s.vel[idx];
            T x = pos_s[4 * i + c];
            if (do_close) v = mul_add_unfused(half_dt, a, v);  // closing kick, nbody.py:214
            if (srow >= 0) {
                s.stage[idx] = x;
                s.stage[n3 + idx] = v;
                s.stage[2 * n3 + idx] = a;
            }
            if (do_open) {
                v = mul_add_unfused(half_dt, a, v);  // opening kick, nbody.py:205
                x = mul_add_unfused(dt, v, x);       // drift, nbody.py:208
                pos_s[4 * i + c] = x;
            }
            s.vel[idx] = v;
        }
        pending = srow;
        __syncthreads();
    }
    if (pending >= 0) flush_stage<T>(g, s, b, pending);
    __syncthreads();  // stage and state are quiescent for whoever comes next
}

template <typename T, bool kZeroEps, int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) ensemble_kernel(const EnsembleArgs g) {
    using V4 = typename Vec4<T>::type;
    extern __shared__ __align__(16) char smem[];
    const int n3 = 3 * g.N;
    SystemSmem<T> s;
    s.pos = reinterpret_cast<V4*>(smem);
    s.vel = reinterpret_cast<T*>(smem + (size_t)g.N * sizeof(V4));
    s.acc = s.vel + n3;
    s.part = s.acc + n3;
    s.stage = s.part + (size_t)g.parts * n3;
    __shared__ int s_claim;

    // Two CTAs share an SM and would run in lock step: both in the force phase (FP64 pipe saturated), then both
    // in the integrate phase (pipe idle).  Delaying every second arrival on an SM by about half a step makes
    // one CTA's integrate phase overlap the other's force phase for the rest of the run.
    if (g.sm_slots != nullptr && g.stagger_ns > 0) {
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int arrival = atomicAdd(g.sm_slots + smid, 1);
            if (arrival & 1) {
                const long long t0 = clock64();
                const long long ticks = (long long)g.stagger_ns * 2;  // ~2 GHz SM clock
                while (clock64() - t0 < ticks) __nanosleep(200);
            }
        }
        __syncthreads();
    }

    // Home systems: CTA c owns systems c, c + grid, ... (whole rounds of the grid) and keeps each one in
    // shared memory from its first step to its last.  Leftover systems (B mod grid of them) are advanced
    // chunk by chunk by whichever CTA claims the next (chunk, system) ticket at one of its own chunk
    // boundaries: it parks its home system in global memory, runs the stolen chunk, and takes its home
    // system back.  Every CTA steals at most `steal_budget` chunks, which spreads the leftover work evenly.
    const int grid = gridDim.x;
    const int home_rounds = g.home_rounds;
    const int b_home = home_rounds * grid;
    const int n_left = g.B - b_home;
    const int n_tickets = n_left * g.n_chunks;
    const int steal_budget = n_left ? (n_tickets + grid - 1) / grid : 0;
    int steals = 0;

    // Claim the next leftover ticket if its predecessor chunk is finished (non-blocking unless `block`).
    auto claim = [&](bool block) -> int {
        __syncthreads();
        if (threadIdx.x == 0) {
            int got = -1;
            if (block) {
                // unconditional draw (atomics pipeline at L2; a CAS loop would serialise 296 CTAs), then wait
                // for the predecessor chunk: it holds a lower ticket, so some running CTA is executing it
                const int t = atomicAdd(g.ticket, 1);
                if (t < n_tickets) {
                    const int chunk = t / n_left, bl = b_home + (t - chunk * n_left);
                    if (chunk > 0) {
                        volatile int* flag = g.progress + bl;
                        while (*flag < chunk) __nanosleep(32);
                    }
                    got = t;
                }
            } else
            for (;;) {
                const int t = *reinterpret_cast<volatile int*>(g.ticket);
                if (t >= n_tickets) break;
                const int chunk = t / n_left, bl = b_home + (t - chunk * n_left);
                const bool ready = chunk == 0 || *reinterpret_cast<volatile int*>(g.progress + bl) >= chunk;
                if (ready) {
                    if (atomicCAS(g.ticket, t, t + 1) == t) { got = t; break; }
                } else if (!block) {
                    break;
                } else {
                    __nanosleep(64);
                }
            }
            if (got >= 0) __threadfence();
            s_claim = got;
        }
        __syncthreads();
        return s_claim;
    };
    auto run_ticket = [&](int t) {
        const int chunk = t / n_left, bl = b_home + (t - chunk * n_left);
        const int k0 = chunk * g.chunk_steps, k1 = min(k0 + g.chunk_steps, g.n_steps);
        load_state<T>(g, bl, s);
        run_steps<T, kZeroEps>(g, bl, k0, k1, s);
        store_state<T>(g, bl, s);
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(g.progress + bl, chunk + 1);
    };

    // Home boundaries (where a CTA may steal) are every kPeekSteps steps, staggered by CTA so that at every
    // step some CTAs are at a boundary and a leftover chunk is picked up as soon as it becomes ready.
    constexpr int kPeekSteps = 8;
    const int phase = blockIdx.x % kPeekSteps;
    for (int round = 0; round < home_rounds; ++round) {
        const int b = blockIdx.x + round * grid;
        load_state<T>(g, b, s);
        int k0 = 0;
        do {
            int k1 = g.n_steps;
            if (n_tickets && steals < steal_budget) {
                const int d = ((k0 - phase) % kPeekSteps + kPeekSteps) % kPeekSteps;
                k1 = min(k0 + (kPeekSteps - d), g.n_steps);
            }
            run_steps<T, kZeroEps>(g, b, k0, k1, s);
            const bool last = k1 >= g.n_steps;
            if (last) store_state<T>(g, b, s);
            if (n_tickets && steals < steal_budget) {
                const int t = claim(false);
                if (t >= 0) {
                    if (!last) store_state<T>(g, b, s);
                    run_ticket(t);
                    ++steals;
                    if (!last) load_state<T>(g, b, s);
                }
            }
            k0 = k1;
        } while (k0 < g.n_steps);
    }
    // whatever leftover work is still unclaimed (short runs with a single chunk, or no home round at all)
    while (n_tickets) {
        const int t = claim(true);
        if (t < 0) break;
        run_ticket(t);
    }
}

constexpr int kEnsembleMaxBodies = 1024;

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
    // two bodies per thread; as many j-parts as fit in 512 threads (at most 8)
    g.rows = ceil_div(N, 2);
    int parts = 512 / g.rows;
    if (parts < 1) parts = 1;
    if (parts > 8) parts = 8;
    if (parts > N) parts = N;
    g.parts = parts;
    const int threads = round_up(g.rows * parts, 32);
    const size_t smem = ensemble_smem_bytes<T>(N, parts);
    const bool zero = !((T)g.eps2 > T(0));
    // up to 512 threads: two CTAs per SM on a 64-register budget; more rows: one CTA per SM
    void (*kern)(const EnsembleArgs);
    if (threads <= 512 && 2 * smem <= 200 * 1024)
        kern = zero ? ensemble_kernel<T, true, 512, 2> : ensemble_kernel<T, false, 512, 2>;
    else
        kern = zero ? ensemble_kernel<T, true, 1024, 1> : ensemble_kernel<T, false, 1024, 1>;
    NB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    NB_REQUIRE(per_sm >= 1, "ensemble kernel does not fit on an SM (N=%d threads=%d smem=%zu)", N, threads, smem);
    const int resident = per_sm * sms;
    const int grid = B < resident ? B : resident;
    // Leftover systems (B mod grid) advance in chunks of 20 steps, each chunk stolen by a home CTA.
    g.ticket = nullptr; g.progress = nullptr; g.chunk_steps = n_steps > 0 ? n_steps : 1; g.n_chunks = 1;
    g.home_rounds = B / grid;
    g.sm_slots = nullptr; g.stagger_ns = 0;
    if (per_sm >= 2 && ws && ws_bytes >= nb_ensemble_workspace_bytes(B)) {
        // half of one shared step: N^2 interactions x 16 FP64 ops on a 64-lane pipe at ~1.9 GHz
        const double step_ns = (double)N * N * 16.0 / (64.0 * 1.9);
        g.stagger_ns = (unsigned)(step_ns > 4.0e6 ? 4.0e6 : step_ns);
        g.sm_slots = static_cast<int*>(ws) + (size_t)B + 1;
        NB_CUDA_OK(cudaMemsetAsync(g.sm_slots, 0, sizeof(int) * 1024, st));
    }
    if (B % grid != 0) {
        NB_REQUIRE(ws && ws_bytes >= nb_ensemble_workspace_bytes(B), "ensemble workspace too small: %zu < %zu",
                   ws_bytes, nb_ensemble_workspace_bytes(B));
        if (B % grid <= grid / 8) {
            // a few leftover systems: 20-step chunks stolen by the home CTAs
            if (n_steps >= 40) {
                g.chunk_steps = 20;
                g.n_chunks = ceil_div(n_steps, 20);
            }
        } else {
            // many: no home systems, every system advances by tickets (~48 per CTA, chunks of >= 8 steps)
            g.home_rounds = 0;
            int steps = ceil_div(n_steps, ceil_div(48 * grid, B));
            if (steps < 8) steps = 8;
            if (n_steps >= 16) {
                g.chunk_steps = steps;
                g.n_chunks = ceil_div(n_steps, steps);
            }
        }
        g.ticket = static_cast<int*>(ws);
        g.progress = g.ticket + 1;
        NB_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(int) * ((size_t)B + 1), st));
    }
    kern<<<grid, threads, smem, st>>>(g);
    return check_launch("ensemble kernel");
}

}  // namespace nb

extern "C" {

int nb_ensemble_max_bodies(void) { return nb::kEnsembleMaxBodies; }

// ticket counter + per-system progress words + per-SM arrival counters
size_t nb_ensemble_workspace_bytes(int B) { return (sizeof(int) * ((size_t)(B > 0 ? B : 0) + 1 + 1024) + 255) / 256 * 256; }

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
