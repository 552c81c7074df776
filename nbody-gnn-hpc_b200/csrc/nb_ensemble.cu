// nb_ensemble.cu -- K3: B independent small systems advanced inside one launch (sm_100a).
//
// Replaces the process pool of the data-generation script, reference
// scripts/generate_data.py:32-58,142-149: there each worker builds one NBodySimulator and calls
// run(n_steps) (src/hpc/nbody.py:220-248), i.e. n_steps x [step() + get_state()].  Here one CTA
// owns one system for a range of steps: positions + G*m live in shared memory, every body has
// `parts` threads that each sum a contiguous j range (partials combined in ascending part order),
// the body's owner thread keeps (x, v, a) in registers, applies the kicks and the drift and
// streams the snapshot rows (x_k, v_k, a_k) straight to HBM in the reference's (T+1, N, 3) layout.
//
// Scheduling.  The grid is persistent: at most (resident CTAs per SM) x (SM count) CTAs.  If the
// ensemble fits (B <= grid) every CTA runs its system start to finish.  Otherwise the run is cut
// into step chunks and CTAs draw (chunk, system) tickets from a global counter, chunk-major; the
// state of a system is handed from chunk to chunk through the in/out state arrays and a per-system
// progress word.  A ticket's predecessor always has a lower ticket number, hence a CTA that is
// already running, so the waits cannot deadlock.  This removes the 300-systems-on-148-SMs tail.
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
    int parts;        // threads per body
    int chunk_steps;  // steps per ticket (dynamic mode)
    int n_chunks;
    int* ticket;      // dynamic mode: global ticket counter, zero on entry
    int* progress;    // dynamic mode: per-system count of finished chunks, zero on entry
};

// Advance system b from step k_begin (state as stored in args.x/v/a) to k_end.
// k_begin == 0 additionally handles the initial acceleration and the initial snapshot.
template <typename T, bool kZeroEps>
__device__ __forceinline__ void advance_system(const EnsembleArgs& g, int b, int k_begin, int k_end,
                                               typename Vec4<T>::type* posm, T* part_acc) {
    using V4 = typename Vec4<T>::type;
    const int N = g.N;
    const int tid = threadIdx.x;
    const int q = tid / N;       // which j-part this thread sums
    const int i = tid - q * N;   // which body
    const bool active = q < g.parts;
    const bool owner = active && q == 0;
    const int jb = active ? (int)(((long)q * N) / g.parts) : 0;
    const int je = active ? (int)(((long)(q + 1) * N) / g.parts) : 0;
    const size_t row = (size_t)N * 3;
    const size_t sbase = (size_t)b * row;
    const T dt = (T)g.dt, half_dt = (T)g.half_dt, eps2 = (T)g.eps2;

    T x[3] = {0, 0, 0}, v[3] = {0, 0, 0}, a[3] = {0, 0, 0};
    __syncthreads();  // previous system's readers are done with posm / part_acc
    if (owner) {
        const char* mb = static_cast<const char*>(g.masses);
        const size_t mi = (size_t)b * g.mass_stride + i;
        const double m = g.masses_are_f32 ? (double)reinterpret_cast<const float*>(mb)[mi]
                                          : reinterpret_cast<const double*>(mb)[mi];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // L2 loads: in ticket mode another SM may have written this state a chunk ago
            x[c] = (T)__ldcg(&g.x[sbase + (size_t)i * 3 + c]);
            v[c] = (T)__ldcg(&g.v[sbase + (size_t)i * 3 + c]);
            a[c] = (T)__ldcg(&g.a[sbase + (size_t)i * 3 + c]);
        }
        V4 p;
        p.x = x[0]; p.y = x[1]; p.z = x[2];
        p.w = (T)(kG * m);  // G * masses[j], nbody.py:57
        posm[i] = p;
    }
    __syncthreads();

    for (int k = k_begin; k <= k_end; ++k) {
        const bool need_force = (k > 0) || g.compute_a0;
        const bool entry_state = (k == k_begin) && (k_begin > 0);  // (x,v,a) of step k_begin were finished by the previous chunk
        if (need_force && !entry_state) {
            T fx = 0, fy = 0, fz = 0;
            if (active) {
                const V4 me = posm[i];
#pragma unroll 4
                for (int j = jb; j < je; ++j) {
                    const V4 pj = posm[j];
                    pair_any<kZeroEps>(me.x, me.y, me.z, pj.x, pj.y, pj.z, pj.w, eps2, fx, fy, fz);
                }
                if (q > 0) {
                    T* pa = part_acc + (size_t)(q - 1) * 3 * N;
                    pa[i] = fx; pa[N + i] = fy; pa[2 * N + i] = fz;
                }
            }
            __syncthreads();
            if (owner) {
                for (int p = 1; p < g.parts; ++p) {
                    const T* pa = part_acc + (size_t)(p - 1) * 3 * N;
                    fx += pa[i]; fy += pa[N + i]; fz += pa[2 * N + i];
                }
                a[0] = fx; a[1] = fy; a[2] = fz;
            }
        }
        if (owner) {
            if (k > 0 && !entry_state) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = mul_add_unfused(half_dt, a[c], v[c]);  // closing kick, nbody.py:214
            }
            // snapshot rows: get_state() before the loop and every save_interval steps, nbody.py:235,240-241
            long srow = -1;
            if (k == 0) {
                if (g.write_initial) srow = g.snap_offset;
            } else if (!entry_state && (k % g.save_interval) == 0) {
                srow = g.snap_offset + (g.write_initial ? 1 : 0) + (k / g.save_interval - 1);
            }
            if (srow >= 0 && g.out_x) {
                const size_t o = ((size_t)b * g.n_snap_total + (size_t)srow) * row + (size_t)i * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    g.out_x[o + c] = (double)x[c];
                    g.out_v[o + c] = (double)v[c];
                    g.out_a[o + c] = (double)a[c];
                }
            }
            if (k < k_end) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v[c] = mul_add_unfused(half_dt, a[c], v[c]);  // opening kick, nbody.py:205
                    x[c] = mul_add_unfused(dt, v[c], x[c]);       // drift, nbody.py:208
                }
                V4 p = posm[i];
                p.x = x[0]; p.y = x[1]; p.z = x[2];
                posm[i] = p;
            }
        }
        __syncthreads();
    }
    if (owner) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g.x[sbase + (size_t)i * 3 + c] = (double)x[c];
            g.v[sbase + (size_t)i * 3 + c] = (double)v[c];
            g.a[sbase + (size_t)i * 3 + c] = (double)a[c];
        }
        __threadfence();  // publish before the progress word is advanced (ticket mode)
    }
}

template <typename T, bool kZeroEps>
__global__ void __launch_bounds__(1024) ensemble_kernel(const EnsembleArgs g) {
    using V4 = typename Vec4<T>::type;
    extern __shared__ __align__(16) char smem[];
    V4* posm = reinterpret_cast<V4*>(smem);
    T* part_acc = reinterpret_cast<T*>(smem + (size_t)g.N * sizeof(V4));
    __shared__ int s_ticket;

    if (g.ticket == nullptr) {  // static: one CTA per system, start to finish
        for (int b = blockIdx.x; b < g.B; b += gridDim.x) advance_system<T, kZeroEps>(g, b, 0, g.n_steps, posm, part_acc);
        return;
    }
    const int n_tickets = g.n_chunks * g.B;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(g.ticket, 1);
        __syncthreads();
        const int t = s_ticket;
        if (t >= n_tickets) break;
        const int chunk = t / g.B;
        const int b = t - chunk * g.B;
        if (chunk > 0) {
            if (threadIdx.x == 0) {
                volatile int* flag = g.progress + b;
                while (*flag < chunk) __nanosleep(64);
                __threadfence();
            }
            __syncthreads();
        }
        const int k0 = chunk * g.chunk_steps;
        const int k1 = min(k0 + g.chunk_steps, g.n_steps);
        advance_system<T, kZeroEps>(g, b, k0, k1, posm, part_acc);
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(g.progress + b, chunk + 1);
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
    // threads per body: fill about 416 threads per CTA (13 warps), at most 8 parts
    int parts = 416 / N;
    if (parts < 1) parts = 1;
    if (parts > 8) parts = 8;
    g.parts = parts;
    const int threads = round_up(N * parts, 32);
    const size_t smem = (size_t)N * sizeof(typename Vec4<T>::type) + (size_t)(parts - 1) * 3 * N * sizeof(T);
    const bool zero = !((T)g.eps2 > T(0));
    auto kern = zero ? ensemble_kernel<T, true> : ensemble_kernel<T, false>;
    NB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    NB_REQUIRE(per_sm >= 1, "ensemble kernel does not fit on an SM (N=%d threads=%d smem=%zu)", N, threads, smem);
    const int resident = per_sm * sms;
    g.ticket = nullptr; g.progress = nullptr; g.chunk_steps = n_steps; g.n_chunks = 1;
    int grid = B < resident ? B : resident;
    if (B > resident && n_steps >= 16) {
        // ticket mode: about a dozen tickets per resident CTA keeps the tail under a few percent
        NB_REQUIRE(ws && ws_bytes >= nb_ensemble_workspace_bytes(B), "ensemble workspace too small: %zu < %zu",
                   ws_bytes, nb_ensemble_workspace_bytes(B));
        int want = ceil_div(12 * resident, B);
        int steps = ceil_div(n_steps, want);
        if (steps < 8) steps = 8;
        g.chunk_steps = steps;
        g.n_chunks = ceil_div(n_steps, steps);
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
