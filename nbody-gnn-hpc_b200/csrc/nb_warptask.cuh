// nb_warptask.cuh -- a WARP-TASK: 32 lanes x kP bodies (a "group") against one j-segment of the position stream,
// streamed global -> shared memory through the warp's OWN ring of 2 KB tiles (1-D bulk TMA issued by lane 0, completion
// on the warp's own mbarriers) and applied with the loops of nb_tiles.cuh -- the bits of K2's segment partials.  Shared
// by the persistent multi-step kernel (nb_persist.cu) and the group-per-CTA step kernel (nb_group.cu).
#pragma once

#include "nb_tiles.cuh"

namespace nb {

constexpr int kPStages = 2;        // ring depth per warp
constexpr int kPTileBytes = 2048;  // 128 float32 bodies / 64 float64 bodies per tile

// mbarrier wait for the warp's own ring; bounded like every other spin in this library
__device__ __forceinline__ bool ring_wait(uint64_t* bar, uint32_t parity) {
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
        if (!done && ++spins > (1u << 24)) return false;
    } while (!done);
    return true;
}

// The bodies of group g that lane `lane` owns: g * 32 * kP + lane + 32 * k, k < kP.
template <typename T, int kP, bool kZeroEps>
struct WarpTask;

template <int kP, bool kZeroEps>
struct WarpTask<float, kP, kZeroEps> {
    // One task: the group's bodies against segment [j0, j1) of stream `cur`; the partial of body li goes to
    // out[c * out_stride + (li - out_first)], c = 0..2 (a segment slab in global memory, or in shared memory).
    static __device__ __forceinline__ bool run(const float* __restrict__ cur, int n, int g, int j0, int j1, float eps2,
                                               float* __restrict__ out, int out_stride, int out_first, char* ring,
                                               uint64_t* bars, uint32_t& tiles_done, int lane) {
        const int b0 = g * 32 * kP;
        float xi[kP], yi[kP], zi[kP];
        float2 ax[kP], ay[kP], az[kP];
#pragma unroll
        for (int k = 0; k < kP; ++k) {
            const int gi = min(b0 + lane + 32 * k, n - 1);
            xi[k] = __ldcg(cur + StreamIO<float>::index(gi, 0));
            yi[k] = __ldcg(cur + StreamIO<float>::index(gi, 1));
            zi[k] = __ldcg(cur + StreamIO<float>::index(gi, 2));
            ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
        }
        const int own_lo = b0, own_hi = min(b0 + 32 * kP, n);
        const char* src = reinterpret_cast<const char*>(cur) + (size_t)j0 * 16;
        const int total = (j1 - j0) * 16;
        const int n_tiles = (total + kPTileBytes - 1) / kPTileBytes;
        // the warp's tiles use the ring slots round robin across tasks and steps: tile number c of the launch sits in
        // slot c % kPStages and completes phase c / kPStages of that slot's barrier
        if (lane == 0) {
            for (int t = 0; t < kPStages && t < n_tiles; ++t) {
                const int slot = (tiles_done + t) % kPStages;
                const int bytes = min(kPTileBytes, total - t * kPTileBytes);
                mbar_arrive_expect_tx(&bars[slot], bytes);
                tma_load_1d(ring + slot * kPTileBytes, src + (size_t)t * kPTileBytes, bytes, &bars[slot]);
            }
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int slot = tiles_done % kPStages;
            if (!ring_wait(&bars[slot], (tiles_done / kPStages) & 1u)) return false;
            ++tiles_done;
            const int bytes = min(kPTileBytes, total - t * kPTileBytes);
            const float4* __restrict__ tile = reinterpret_cast<const float4*>(ring + slot * kPTileBytes);
            const int n_pairs = bytes >> 5;
            const int jt_lo = j0 + t * (kPTileBytes / 16), jt_hi = jt_lo + 2 * n_pairs;
            if (jt_lo < own_hi && own_lo < jt_hi) f32_pairs<kP, true>(tile, n_pairs, xi, yi, zi, ax, ay, az, eps2);
            else f32_pairs<kP, false>(tile, n_pairs, xi, yi, zi, ax, ay, az, eps2);
            __syncwarp();  // every lane is done with this slot before it is refilled
            const int nt = t + kPStages;
            if (lane == 0 && nt < n_tiles) {
                const int nbytes = min(kPTileBytes, total - nt * kPTileBytes);
                mbar_arrive_expect_tx(&bars[slot], nbytes);
                tma_load_1d(ring + slot * kPTileBytes, src + (size_t)nt * kPTileBytes, nbytes, &bars[slot]);
            }
        }
#pragma unroll
        for (int k = 0; k < kP; ++k) {
            const int li = b0 + lane + 32 * k;
            if (li < n) {
                const int o = li - out_first;
                out[o] = ax[k].x + ax[k].y;  // even-j lane + odd-j lane, as K2
                out[(size_t)out_stride + o] = ay[k].x + ay[k].y;
                out[(size_t)2 * out_stride + o] = az[k].x + az[k].y;
            }
        }
        return true;
    }
};

template <int kP, bool kZeroEps>
struct WarpTask<double, kP, kZeroEps> {
    static __device__ __forceinline__ bool run(const double* __restrict__ cur, int n, int g, int j0, int j1, double eps2,
                                               double* __restrict__ out, int out_stride, int out_first, char* ring,
                                               uint64_t* bars, uint32_t& tiles_done, int lane) {
        const int b0 = g * 32 * kP;
        double xi[kP], yi[kP], zi[kP], ax[kP], ay[kP], az[kP];
#pragma unroll
        for (int k = 0; k < kP; ++k) {
            const int gi = min(b0 + lane + 32 * k, n - 1);
            xi[k] = __ldcg(cur + StreamIO<double>::index(gi, 0));
            yi[k] = __ldcg(cur + StreamIO<double>::index(gi, 1));
            zi[k] = __ldcg(cur + StreamIO<double>::index(gi, 2));
            ax[k] = ay[k] = az[k] = 0.0;
        }
        const char* src = reinterpret_cast<const char*>(cur) + (size_t)j0 * 32;
        const int total = (j1 - j0) * 32;
        const int n_tiles = (total + kPTileBytes - 1) / kPTileBytes;
        // the warp's tiles use the ring slots round robin across tasks and steps: tile number c of the launch sits in
        // slot c % kPStages and completes phase c / kPStages of that slot's barrier
        if (lane == 0) {
            for (int t = 0; t < kPStages && t < n_tiles; ++t) {
                const int slot = (tiles_done + t) % kPStages;
                const int bytes = min(kPTileBytes, total - t * kPTileBytes);
                mbar_arrive_expect_tx(&bars[slot], bytes);
                tma_load_1d(ring + slot * kPTileBytes, src + (size_t)t * kPTileBytes, bytes, &bars[slot]);
            }
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int slot = tiles_done % kPStages;
            if (!ring_wait(&bars[slot], (tiles_done / kPStages) & 1u)) return false;
            ++tiles_done;
            const int bytes = min(kPTileBytes, total - t * kPTileBytes);
            f64_bodies<kP, kZeroEps>(reinterpret_cast<const double2*>(ring + slot * kPTileBytes), bytes >> 5, xi, yi, zi,
                                     ax, ay, az, eps2);
            __syncwarp();
            const int nt = t + kPStages;
            if (lane == 0 && nt < n_tiles) {
                const int nbytes = min(kPTileBytes, total - nt * kPTileBytes);
                mbar_arrive_expect_tx(&bars[slot], nbytes);
                tma_load_1d(ring + slot * kPTileBytes, src + (size_t)nt * kPTileBytes, nbytes, &bars[slot]);
            }
        }
#pragma unroll
        for (int k = 0; k < kP; ++k) {
            const int li = b0 + lane + 32 * k;
            if (li < n) {
                const int o = li - out_first;
                out[o] = ax[k];
                out[(size_t)out_stride + o] = ay[k];
                out[(size_t)2 * out_stride + o] = az[k];
            }
        }
        return true;
    }
};

}  // namespace nb
