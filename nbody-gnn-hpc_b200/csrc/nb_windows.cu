// nb_windows.cu -- K5: sliding-window training samples straight from the snapshot stacks (sm_100a).
//
// Replaces the sample loop of create_training_dataset, reference src/hpc/checkpoint.py:362-384:
//     for i in range(0, n_steps - L, stride):
//         input  = concat(positions[i:i+L], velocities[i:i+L], axis=-1).astype(float32)   # (L, N, 6)
//         target = concat(positions[i+L],   velocities[i+L],   axis=-1).astype(float32)   # (N, 6)
// for every trajectory of an ensemble whose float64 stacks (B, rows, N, 3) are still in HBM (the
// output of K3): inputs (B*S, L, N, 6) and targets (B*S, N, 6) float32, trajectory-major, in the
// layout the reference's dataset file holds (METHODOLOGY.md:182-187).  Every state is written L+1
// times, so this is pure data movement: 24*N bytes read per state, 24*N*(L+1) bytes written per
// sample -- HBM-write bound.
//
// One "state vector" is [x y z vx vy vz] x N float32 = 24*N bytes, and a sample's input is L
// CONSECUTIVE state vectors.  The B*S samples are cut into one contiguous range per CTA (several CTAs
// per SM); a CTA streams the states of its range ONCE, in time order, through a shared-memory ring of
// float32 state vectors -- a batch of K states per iteration: all threads load the float64
// coordinates (independent 8-byte loads, ~19 in flight per thread), convert, and store them into the
// ring -- and every sample whose last state has just arrived leaves as bulk copies shared -> global
// (cp.async.bulk.global.shared::cta, SASS UBLKCP): L*24*N contiguous bytes for the input (two pieces
// when the window wraps around the ring) and 24*N for the target.  No per-element store
// instructions; the copies of one batch drain while the next batch is being loaded
// (cp.async.bulk.wait_group.read 1 keeps exactly one batch of copies in flight, and the ring holds
// 2K + L + 1 states so that nothing still being read is overwritten).  24*N must be a multiple of 16
// (N even) for the bulk path; other N, and states too large for the ring, take the element-wise kernel.
// The samples are cut into one range per CTA, three CTAs per SM at N = 200.
#include <stdlib.h>

#include "nb_common.cuh"

namespace nb {

struct WindowArgs {
    const double* pos;  // (B, rows, N, 3)
    const double* vel;
    float* inputs;      // (B*S, L, N, 6)
    float* targets;     // (B*S, N, 6)
    int B, rows, N, L, stride, S;
    int ring;           // ring size in states
    int batch;          // K: states per iteration
};

__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}

__global__ void __launch_bounds__(256) window_stream_kernel(const WindowArgs g) {
    extern __shared__ __align__(128) float ring[];  // g.ring x N x 6
    const int n3 = 3 * g.N, n6 = 6 * g.N;
    const uint32_t state_bytes = (uint32_t)g.N * 24u;
    const long Q = (long)g.B * g.S;
    long q = (long)blockIdx.x * Q / gridDim.x;             // this CTA's samples [q, q_end)
    const long q_end = (long)(blockIdx.x + 1) * Q / gridDim.x;

    while (q < q_end) {
        // a run: consecutive samples [s_a, s_b) of one trajectory
        const int b = (int)(q / g.S);
        const int s_a = (int)(q - (long)b * g.S);
        const int s_b = (int)min((long)g.S, s_a + (q_end - q));
        const int i_a = s_a * g.stride;                    // first state of the run
        const int i_b = (s_b - 1) * g.stride + g.L;        // last state of the run (target of the last sample)
        const size_t base = ((size_t)b * g.rows + i_a) * n3;
        int s_next = s_a;                                  // next sample to leave
        for (int i0 = i_a; i0 <= i_b; i0 += g.batch) {
            const int k_states = min(g.batch, i_b - i0 + 1);
            // every copy but those of the previous batch has finished reading the ring
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            const int total = k_states * n3;
            const size_t off = (size_t)(i0 - i_a) * n3;
#pragma unroll 4
            for (int e = threadIdx.x; e < total; e += blockDim.x) {
                const double px = __ldg(g.pos + base + off + e);
                const double vx = __ldg(g.vel + base + off + e);
                const int st = e / n3, j = e - st * n3;
                const int body = j / 3, c = j - 3 * body;
                float* dst = ring + (size_t)((i0 - i_a + st) % g.ring) * n6 + body * 6 + c;
                dst[0] = (float)px;  // .astype(np.float32): round to nearest
                dst[3] = (float)vx;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the copies read through the async proxy
            __syncthreads();
            // samples completed by this batch: s*stride + L <= i0 + k_states - 1
            const int i_last = i0 + k_states - 1;
            int s_done = (i_last - g.L) >= 0 ? (i_last - g.L) / g.stride + 1 : 0;   // samples with index < s_done are complete
            s_done = min(s_done, s_b);
            for (int s = s_next + threadIdx.x; s < s_done; s += blockDim.x) {
                const size_t sample = (size_t)b * g.S + s;
                const int first = (s * g.stride - i_a) % g.ring;                   // ring slot of the window's first state
                NB_CHECK(first >= 0 && first < g.ring && g.L < g.ring);
                const int run1 = min(g.L, g.ring - first);                          // states before the ring wraps
                float* in = g.inputs + sample * (size_t)g.L * n6;
                bulk_store(in, ring + (size_t)first * n6, (uint32_t)run1 * state_bytes);
                if (run1 < g.L) bulk_store(in + (size_t)run1 * n6, ring, (uint32_t)(g.L - run1) * state_bytes);
                bulk_store(g.targets + sample * n6, ring + (size_t)((first + g.L) % g.ring) * n6, state_bytes);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            s_next = max(s_next, s_done);
        }
        // the ring is reused from slot 0 by the next run
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        q += s_b - s_a;
    }
}

// any N: one thread per (sample, l, body) of inputs / (sample, body) of targets, six floats each
__global__ void __launch_bounds__(256) window_elementwise_kernel(const WindowArgs g) {
    const size_t per_sample = (size_t)(g.L + 1) * g.N;
    const size_t total = (size_t)g.B * g.S * per_sample;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t sample = e / per_sample;
        const int r = (int)(e - sample * per_sample);
        const int l = r / g.N, body = r - l * g.N;
        const int b = (int)(sample / g.S), s = (int)(sample - (size_t)b * g.S);
        const size_t src = (((size_t)b * g.rows + (size_t)s * g.stride + l) * g.N + body) * 3;
        float* dst = l < g.L ? g.inputs + ((sample * g.L + l) * g.N + body) * 6 : g.targets + (sample * g.N + body) * 6;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dst[c] = (float)__ldg(g.pos + src + c);
            dst[3 + c] = (float)__ldg(g.vel + src + c);
        }
    }
}

}  // namespace nb

extern "C" {

// Samples per trajectory: len(range(0, n_states - L, stride)), reference checkpoint.py:333,365
int nb_window_count(int n_states, int sequence_length, int stride) {
    if (sequence_length < 1 || stride < 1 || n_states - sequence_length <= 0) return 0;
    return (n_states - sequence_length + stride - 1) / stride;
}

int nb_window_gather_f32(const double* pos, const double* vel, int B, int rows, int N, int n_states, int sequence_length,
                         int stride, float* inputs, float* targets, nb_stream_t s) {
    NB_REQUIRE(pos && vel && inputs && targets, "null pointer argument");
    NB_REQUIRE(B > 0 && N > 0 && rows > 0 && n_states > 0 && n_states <= rows,
               "need B, N > 0 and 0 < n_states <= rows (got B=%d N=%d n_states=%d rows=%d)", B, N, n_states, rows);
    NB_REQUIRE(sequence_length >= 1 && stride >= 1, "sequence_length and stride must be >= 1");
    cudaStream_t st = (cudaStream_t)s;
    nb::WindowArgs g;
    g.pos = pos; g.vel = vel; g.inputs = inputs; g.targets = targets;
    g.B = B; g.rows = rows; g.N = N; g.L = sequence_length; g.stride = stride;
    g.S = nb_window_count(n_states, sequence_length, stride);
    if (g.S == 0) return NB_OK;  // "No samples could be created" is the caller's error to raise
    int dev = 0, sms = 0, smem_max = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t state_bytes = (size_t)N * 24;
    // As many CTAs per SM (at most 4) as still leave each a ring of 2K + L + 1 states with K >= 1: three 72 KB rings
    // at N = 200, L = 10.  Measured at the data-generation shape (300 x 401 x 200): 1 / 2 / 3 CTAs per SM ->
    // 2.15 / 1.45 / 1.22 ms; several small rings keep more loads and more copies in flight than one large ring.
    int per_sm = 4;
    long batch = 0;
    if (const char* e = getenv("NB_WINDOW_CTAS_PER_SM")) per_sm = atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : 4;  // tuning aid
    for (; per_sm >= 1; --per_sm) {
        const size_t budget = ((size_t)smem_max - 4096) / per_sm - 1024;
        batch = ((long)(budget / state_bytes) - sequence_length - 1) / 2;
        if (batch >= 1 || getenv("NB_WINDOW_CTAS_PER_SM")) break;
    }
    if (per_sm < 1) per_sm = 1;
    if (batch > 8) batch = 8;
    const bool bulk = (state_bytes % 16 == 0) && batch >= 1 && (size_t)sequence_length * state_bytes < (1u << 30);
    if (bulk) {
        g.batch = (int)batch;
        g.ring = 2 * g.batch + sequence_length + 1;
        const size_t smem = (size_t)g.ring * state_bytes;
        NB_CUDA_OK(cudaFuncSetAttribute(nb::window_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long Q = (long)B * g.S;
        const int grid = (int)(Q < (long)per_sm * sms ? Q : (long)per_sm * sms);
        nb::window_stream_kernel<<<grid, 256, smem, st>>>(g);
    } else {
        g.ring = g.batch = 0;
        const size_t work = (size_t)B * g.S * (sequence_length + 1) * N;
        const size_t want = (work + 255) / 256;
        const int blocks = (int)(want < (size_t)sms * 32 ? want : (size_t)sms * 32);
        nb::window_elementwise_kernel<<<blocks, 256, 0, st>>>(g);
    }
    return nb::check_launch("window gather kernel");
}

}  // extern "C"
