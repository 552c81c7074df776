// nb_energy.cu -- K4: kinetic and potential energy of a slab of bodies, float64 throughout (sm_100a).
//
// Replaces compute_total_energy, reference src/hpc/nbody.py:101-130 (single-threaded there) and the
// per-step energy of src/utils/metrics.py:62-109 (O(N^2) memory there).
//   K = sum_i 1/2 m_i |v_i|^2                                   nbody.py:116-118
//   U = - sum_{i<j} G m_i m_j / sqrt(r_ij^2 + eps^2)             nbody.py:122-128
// evaluated as U = -1/2 sum_i m_i sum_{j != i} G m_j / sqrt(...) so that rows [i0, i0+n_i) can be
// owned by one rank; ranks add their (K, U).  grid = (blocks of 128 rows, j-splits): small slabs are cut along j as
// well so that the grid still covers the SMs (N = 16,384: 128 row blocks x 5 splits; 0.99 -> 0.3 ms).  Reduction
// order is fixed: thread-sequential over j, warp shuffle tree, warp partials in warp order, (block, split) partials
// in that order.
#include "nb_common.cuh"

namespace nb {

constexpr int kEnergyBlock = 128;
constexpr int kEnergyTile = 128;
constexpr int kEnergyMaxSplits = 64;

__device__ __forceinline__ double rsqrt_f64(double r2) {
    const double y0 = rsqrt_seed(r2);
    const double e = fma(-(r2 * y0), y0, 1.0);
    return fma(y0, e * fma(0.375, e, 0.5), y0);  // y0 * (1 + e/2 + 3/8 e^2)
}

__global__ void __launch_bounds__(kEnergyBlock)
energy_kernel(const double* __restrict__ pos, const double* __restrict__ vel, const void* __restrict__ masses,
              int masses_are_f32, int n, int i0, int n_i, double eps2, int split_len,
              double* __restrict__ block_ku) {
    __shared__ double4 tile[kEnergyTile];
    __shared__ double warp_k[kEnergyBlock / 32], warp_u[kEnergyBlock / 32];
    const int li = blockIdx.x * kEnergyBlock + threadIdx.x;
    const bool valid = li < n_i;
    const int gi = i0 + (valid ? li : 0);
    auto mass = [&](int idx) {
        return masses_are_f32 ? (double)static_cast<const float*>(masses)[idx] : static_cast<const double*>(masses)[idx];
    };
    const double xi = pos[(size_t)gi * 3 + 0], yi = pos[(size_t)gi * 3 + 1], zi = pos[(size_t)gi * 3 + 2];
    double phi = 0.0;
    const int j_begin = blockIdx.y * split_len, j_end = min(n, j_begin + split_len);  // split_len % kEnergyTile == 0
    for (int j0 = j_begin; j0 < j_end; j0 += kEnergyTile) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (j < n) tile[threadIdx.x] = make_double4(pos[(size_t)j * 3], pos[(size_t)j * 3 + 1], pos[(size_t)j * 3 + 2], kG * mass(j));
        else tile[threadIdx.x] = make_double4(0.0, 0.0, 0.0, 0.0);
        __syncthreads();
        const int cnt = min(kEnergyTile, j_end - j0);
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const double4 p = tile[t];
            const double dx = p.x - xi, dy = p.y - yi, dz = p.z - zi;
            const double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, eps2)));
            const double term = p.w * rsqrt_f64(r2);
            phi += (j0 + t != gi && r2 > 0.0) ? term : 0.0;
        }
    }
    double k = 0.0, u = 0.0;
    if (valid) {
        const double m = mass(gi);
        const double vx = vel[(size_t)gi * 3 + 0], vy = vel[(size_t)gi * 3 + 1], vz = vel[(size_t)gi * 3 + 2];
        if (blockIdx.y == 0) k = 0.5 * m * (vx * vx + vy * vy + vz * vz);
        u = -0.5 * m * phi;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        k += __shfl_down_sync(0xffffffffu, k, off);
        u += __shfl_down_sync(0xffffffffu, u, off);
    }
    if ((threadIdx.x & 31) == 0) {
        warp_k[threadIdx.x >> 5] = k;
        warp_u[threadIdx.x >> 5] = u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bk = 0.0, bu = 0.0;
        for (int w = 0; w < kEnergyBlock / 32; ++w) {
            bk += warp_k[w];
            bu += warp_u[w];
        }
        const size_t slot = (size_t)blockIdx.x * gridDim.y + blockIdx.y;
        block_ku[2 * slot + 0] = bk;
        block_ku[2 * slot + 1] = bu;
    }
}

__global__ void energy_final_kernel(const double* __restrict__ block_ku, int n_blocks, double* __restrict__ out_ku) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double k = 0.0, u = 0.0;
        for (int b = 0; b < n_blocks; ++b) {
            k += block_ku[2 * b];
            u += block_ku[2 * b + 1];
        }
        out_ku[0] = k;
        out_ku[1] = u;
    }
}

}  // namespace nb

extern "C" {

size_t nb_energy_workspace_bytes(int n, int n_i) {
    (void)n;
    const size_t blocks = (size_t)nb::ceil_div(n_i > 0 ? n_i : 1, nb::kEnergyBlock);
    const size_t splits = blocks >= 1024 ? 1 : nb::kEnergyMaxSplits;  // upper bound of what the launch may choose
    return (blocks * splits * 2 * sizeof(double) + 255) / 256 * 256;
}

int nb_energy_f64(const double* pos, const double* vel, const void* masses, int masses_are_f32, int n, int i0, int n_i,
                  double softening, double* out_ku, void* ws, size_t ws_bytes, nb_stream_t s) {
    NB_REQUIRE(pos && vel && masses && out_ku && ws, "null pointer argument");
    NB_REQUIRE(n > 0 && i0 >= 0 && n_i > 0 && i0 + n_i <= n, "slab [%d, %d) outside system of %d bodies", i0, i0 + n_i, n);
    NB_REQUIRE(ws_bytes >= nb_energy_workspace_bytes(n, n_i), "energy workspace too small");
    const int blocks = nb::ceil_div(n_i, nb::kEnergyBlock);
    cudaStream_t st = (cudaStream_t)s;
    // j-splits: enough CTAs for ~4 per SM, each split a whole number of tiles
    int dev = 0, sms = 0;
    NB_CUDA_OK(cudaGetDevice(&dev));
    NB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int splits = blocks >= 1024 ? 1 : nb::ceil_div(4 * sms, blocks);
    const int tiles = nb::ceil_div(n, nb::kEnergyTile);
    if (splits > nb::kEnergyMaxSplits) splits = nb::kEnergyMaxSplits;
    if (splits > tiles) splits = tiles;
    const int split_len = nb::ceil_div(tiles, splits) * nb::kEnergyTile;
    splits = nb::ceil_div(n, split_len);
    nb::energy_kernel<<<dim3(blocks, splits), nb::kEnergyBlock, 0, st>>>(pos, vel, masses, masses_are_f32, n, i0, n_i,
                                                                       softening * softening, split_len,
                                                                       static_cast<double*>(ws));
    if (int rc = nb::check_launch("energy kernel")) return rc;
    nb::energy_final_kernel<<<1, 32, 0, st>>>(static_cast<const double*>(ws), blocks * splits, out_ku);
    return nb::check_launch("energy final kernel");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// K4b -- energy and momentum of EVERY snapshot of a stack of trajectories.
//
// Replaces compute_energy_error (reference src/utils/metrics.py:62-109: a Python loop over the stored steps with an
// N x N x 3 temporary per step) and compute_momentum_error (:112-137) for B trajectories x S snapshots at once:
//   K = 1/2 sum_i m_i |v_i|^2                                  metrics.py:86
//   U = -1/2 G sum_{i != j} m_i m_j / sqrt(|x_i - x_j|^2 + eps^2)   metrics.py:90-102
//   p = sum_i m_i v_i                                           metrics.py:131
// One CTA per (trajectory, snapshot): the snapshot's positions and masses go to shared memory once, thread i sums the
// pair terms of bodies i+1 .. i+N/2 (cyclically), so every unordered pair is evaluated ONCE and every thread does the
// same work -- half the arithmetic of the row-wise form; for even N the pairs at distance N/2 would be met from both
// ends and are taken by the lower-numbered end only.  Reduction order is fixed (thread-sequential, shuffle tree, warp order).
// ------------------------------------------------------------------------------------------------------------------
namespace nb {

constexpr int kSnapEnergyMaxBodies = 4096;  // 4 doubles + 1 float per body in shared memory (144 KB)

// kF32Product: the masses are float32 and the reference forms m_i * m_j in float32 (np.outer of a float32 array,
// metrics.py:82): the product is rounded the same way here, so the energies agree to float64 rounding, not 1e-8.
template <bool kF32Product>
__global__ void __launch_bounds__(256)
snapshot_energy_kernel(const double* __restrict__ pos, const double* __restrict__ vel,
                       const void* __restrict__ masses, int masses_are_f32, int mass_stride, int S, int N, double G,
                       double eps2, double* __restrict__ out) {
    extern __shared__ __align__(16) double4 body[];  // x, y, z, m; behind them (kF32Product) the masses as float
    __shared__ double red[5][8];
    float* mf = reinterpret_cast<float*>(body + N);
    const int snap = blockIdx.x;                  // b * S + s
    const int b = snap / S;
    const double* x = pos + (size_t)snap * N * 3;
    const double* v = vel + (size_t)snap * N * 3;
    auto mass = [&](int i) {
        const size_t mi = (size_t)b * mass_stride + i;
        return masses_are_f32 ? (double)static_cast<const float*>(masses)[mi] : static_cast<const double*>(masses)[mi];
    };
    const bool potential = pos != nullptr;  // momentum / kinetic energy only (compute_momentum_error) when null
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double m = mass(i);
        body[i] = potential ? make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], m) : make_double4(0., 0., 0., m);
        if (kF32Product) mf[i] = (float)m;
    }
    __syncthreads();
    double k = 0.0, u = 0.0, px = 0.0, py = 0.0, pz = 0.0;
    const int half = N / 2;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double4 me = body[i];
        const float mfi = kF32Product ? mf[i] : 0.f;
        const double vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
        k += 0.5 * me.w * (vx * vx + vy * vy + vz * vz);
        px += me.w * vx;
        py += me.w * vy;
        pz += me.w * vz;
        // partners i+1 .. i+cnt (mod N); for even N the pair at distance N/2 is shared with the partner and the lower
        // index takes it.  Two straight runs -- up to the end of the array, then from its start -- so the loops have
        // plain increments and unroll.
        const int cnt = !potential ? 0 : (N % 2 == 0 && i >= half) ? half - 1 : half;
        double phi = 0.0;
        auto run = [&](int j0, int j1) {
#pragma unroll 4
            for (int j = j0; j < j1; ++j) {
                const double4 p = body[j];
                const double dx = p.x - me.x, dy = p.y - me.y, dz = p.z - me.z;
                const double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, eps2)));
                // float32 masses: the product is formed in float32 like the reference's np.outer (one FMUL, one
                // conversion: conversions share the MUFU pipe)
                const double mm = kF32Product ? (double)__fmul_rn(mfi, mf[j]) : p.w;
                phi += r2 > 0.0 ? mm * rsqrt_f64(r2) : 0.0;
            }
        };
        NB_CHECK(cnt >= 0 && cnt <= half && i < N);
        const int first_end = min(N, i + 1 + cnt);
        run(i + 1, first_end);
        run(0, cnt - (first_end - (i + 1)));
        u -= kF32Product ? G * phi : G * me.w * phi;
    }
    double vals[5] = {k, u, px, py, pz};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        double s = vals[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        out[(size_t)snap * 5 + threadIdx.x] = s;
    }
}

}  // namespace nb

extern "C" {

int nb_snapshot_energy_max_bodies(void) { return nb::kSnapEnergyMaxBodies; }

int nb_snapshot_energy_f64(const double* pos, const double* vel, const void* masses, int masses_are_f32,
                           int mass_stride, int B, int S, int N, double G, double softening, double* out,
                           nb_stream_t s) {
    NB_REQUIRE(vel && masses && out, "null pointer argument");  // pos may be null: K and p only, U = 0
    NB_REQUIRE(B > 0 && S > 0 && N > 0 && N <= nb::kSnapEnergyMaxBodies, "need B, S > 0 and 0 < N <= %d (got %d %d %d)",
               nb::kSnapEnergyMaxBodies, B, S, N);
    NB_REQUIRE(mass_stride == 0 || mass_stride == N, "mass_stride must be 0 (shared) or N");
    NB_REQUIRE((long long)B * S <= 0x7fffffffLL, "too many snapshots");
    const size_t smem = (size_t)N * (sizeof(double4) + sizeof(float));
    auto kern = masses_are_f32 ? nb::snapshot_energy_kernel<true> : nb::snapshot_energy_kernel<false>;
    NB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = nb::round_up(N < 256 ? N : 256, 32);
    kern<<<B * S, threads, smem, (cudaStream_t)s>>>(pos, vel, masses, masses_are_f32, mass_stride, S, N, G,
                                                    softening * softening, out);
    return nb::check_launch("snapshot energy kernel");
}

}  // extern "C"
