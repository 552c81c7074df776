// nb_force.cu -- K1 (tiled i-j force) and K2 (force pass fused with the leapfrog) for one
// system held in stream layout, float32 and float64, for sm_100a.
//
// Replaces compute_accelerations_direct (reference src/hpc/nbody.py:22-66) and
// NBodySimulator.step / the loop of run (src/hpc/nbody.py:202-218, 237-241).
//
// Decomposition.  grid = (i-tiles, j-segments).  A CTA owns kBlock*kP bodies i (kP per thread, in
// registers) and one j-segment; the segment is streamed global -> shared memory as 4 KB tiles by
// 1-D bulk TMA copies (cp.async.bulk, completion on an mbarrier, 4-deep ring), every thread reads
// each j (pair) with broadcast LDS.128 and applies it to its kP bodies.  Segment partials go to
// the workspace; the CTA that arrives LAST at its i-tile's counter (tile_epilogue) adds them in
// ascending segment order and, for K2, applies the closing kick, writes the snapshot, applies the
// next opening kick and drift, and writes the new positions into the other stream buffer -- force
// pass and leapfrog are one launch and positions never leave HBM between steps.  In the sharded
// mode (nb_step_peer_*) the same epilogue stores the drifted records straight into every rank's
// next stream over NVLink, tile by tile while other tiles still compute, and the last tile of the
// launch exchanges arrival words with the peers: force, leapfrog and collective are ONE kernel.
//
// float32 inner loop: two j bodies per instruction through the packed f32x2 pipe
// (FADD2/FFMA2/FMUL2), 12 FMA-pipe lane-operations + 1 MUFU.RSQ per interaction.
// float64 inner loop: pair_f64() in nb_common.cuh, 16 FP64-pipe operations + 1 MUFU.RSQ64H.
#include <stdlib.h>

#include <type_traits>

#include "nb_tiles.cuh"

namespace nb {

constexpr int kStages = 4;
constexpr int kTileBytes = 4096;

// Stream the byte range [src, src + total_bytes) through the shared-memory ring, calling
// consume(tile_ptr, tile_bytes, tile_index) on every tile by all threads of the CTA.
template <class Consume>
__device__ __forceinline__ void stream_tiles(const char* __restrict__ src, int total_bytes, char* ring,
                                             uint64_t* bars, Consume&& consume) {
    const int n_tiles = (total_bytes + kTileBytes - 1) / kTileBytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 0; t < kStages && t < n_tiles; ++t) {
            const int bytes = min(kTileBytes, total_bytes - t * kTileBytes);
            mbar_arrive_expect_tx(&bars[t], bytes);
            tma_load_1d(ring + t * kTileBytes, src + (size_t)t * kTileBytes, bytes, &bars[t]);
        }
    }
    for (int t = 0; t < n_tiles; ++t) {
        const int slot = t % kStages;
        mbar_wait(&bars[slot], (t / kStages) & 1);
        const int bytes = min(kTileBytes, total_bytes - t * kTileBytes);
        NB_CHECK(bytes > 0 && bytes % 32 == 0 && slot >= 0 && slot < kStages);
        NB_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        consume(ring + slot * kTileBytes, bytes, t);
        __syncthreads();  // every thread is done with this slot before it is refilled
        const int nt = t + kStages;
        if (threadIdx.x == 0 && nt < n_tiles) {
            const int nbytes = min(kTileBytes, total_bytes - nt * kTileBytes);
            mbar_arrive_expect_tx(&bars[slot], nbytes);
            tma_load_1d(ring + slot * kTileBytes, src + (size_t)nt * kTileBytes, nbytes, &bars[slot]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cross-GPU ordering for the sharded mode (nb_step_peer_*): arrival words in peer-visible memory
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;

struct PeerWait {          // force pass: do not read the stream before every rank has published wait_seq
    const uint32_t* flags;  // local array, one word per rank (null: no wait)
    int n_ranks;
    uint32_t seq;
    unsigned long long timeout_ns;  // a rank that has not arrived after this long is declared lost
    int* error;             // the workspace's error word (null outside the sharded mode): sticky, checked by the host
};
struct PeerTargets {       // finish pass: where the new slab goes and whom to tell
    void* next[kMaxPeers];       // every rank's next-stream buffer (own included)
    uint32_t* flags[kMaxPeers];  // every rank's flag array
    int n_ranks, my_rank;
    uint32_t seq;
    unsigned long long timeout_ns;
    int* error;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Error word: phase in the low byte, the rank that never arrived above it.  First failure wins.
__device__ __forceinline__ void peer_lost(int* error, int phase, int rank) {
    atomicCAS(error, 0, phase | (rank << 8));
    __threadfence_system();
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Programmatic dependent launch: the step kernels of a run are launched back to back on one stream.  Each lets
// its successor be scheduled at once (launch_dependents) and, before touching anything a predecessor wrote --
// streams, velocities, partials, counters --, waits for the predecessor grid to have completed and flushed
// (wait).  The successor's launch latency and prologue then overlap this kernel's tail: ~3 us per step, which is
// a quarter of a step at N = 2,048 and 3% at N = 16,384.
__device__ __forceinline__ void pdl_prologue() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Called when a CTA's force loop is done: from here on its successor's CTAs may be placed on the SMs.  (Releasing
// them at kernel entry packs the waiting successor CTAs onto whichever SMs have room first and unbalances the next
// step: 31 -> 58 us at N = 4,096 float64, measured.)
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Called by every thread at kernel entry; one thread per rank polls (bounded by timeout_ns) until all ranks have
// arrived.  Returns false -- and the CTA leaves the kernel without touching anything -- when a rank never arrived
// (the workspace's error word then says which; nb_step_status reports it) or when an earlier launch on this
// workspace already failed: a lost peer is an error the host sees, never a step computed from stale positions.
__device__ __forceinline__ bool peer_wait(const PeerWait& w) {
    if (w.error == nullptr) return true;  // not the sharded mode
    int bad = 0;
    if (threadIdx.x == 0) bad = *reinterpret_cast<volatile int*>(w.error) != 0;
    if (w.flags != nullptr && threadIdx.x < w.n_ranks) {  // one polling thread per rank: the loads overlap
        const unsigned long long t0 = global_ns();
        // sequence numbers wrap: compare as signed distance
        while ((int32_t)(ld_acquire_sys(w.flags + threadIdx.x) - w.seq) < 0) {
            __nanosleep(100);
            if (global_ns() - t0 > w.timeout_ns) {
                peer_lost(w.error, NB_PEER_LOST_BEFORE_FORCE, (int)threadIdx.x);
                bad = 1;
                break;
            }
        }
    }
    if (__syncthreads_or(bad)) return false;
    // the stream is read next through the async (TMA) proxy, issued by thread 0
    if (threadIdx.x == 0) asm volatile("fence.proxy.async;" ::: "memory");
    return true;
}

// ------------------------------------------------------------------------------------------------
// float32 force kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float f32_stream_coord(const float* __restrict__ stream, int body, int c) {
    return stream[(size_t)(body >> 1) * 8 + 2 * c + (body & 1)];
}

// ------------------------------------------------------------------------------------------------
// stream access and the per-body finish (ordered segment reduction + leapfrog)
// ------------------------------------------------------------------------------------------------
enum { kEpiNone = 0, kEpiAccel = 1, kEpiStep = 2, kEpiStepPeer = 3 };

// What the LAST CTA to finish an i-tile does with the tile's bodies, inside the force kernel itself.
template <typename T>
struct Epilogue {
    int mode;           // kEpiNone: partials only; kEpiAccel: acc = sum; kEpiStep: + kicks, snapshot, drift;
                        // kEpiStepPeer: the drifted records go to every rank's stream + arrival words
    int n_seg;
    int* tile_counter;  // one word per i-tile, zero before the launch, left zero
    int* done_counter;  // kEpiStepPeer: finished i-tiles of this launch, zero before, left zero
    PeerTargets peers;  // kEpiStepPeer only
    const T* cur;
    T* next;
    T* vel;
    T* acc;
    T dt, half_dt;
    int flags;
    double* sp;
    double* sv;
    double* sa;
};

// Body li of the slab: add its segment partials in ascending order, then (kEpiStep) the closing kick, the
// snapshot row, and with NB_STEP_CONTINUE the next opening kick and the drift into the other stream.
template <typename T>
__device__ __forceinline__ void finish_body(const Epilogue<T>& e, const T* partial, int i0, int n_i, int li,
                                            T x_out[3]) {
    NB_CHECK(li >= 0 && li < n_i && e.n_seg >= 1 && e.n_seg <= 64);
    T a[3] = {T(0), T(0), T(0)};
    for (int s = 0; s < e.n_seg; ++s) {
        const T* p = partial + (size_t)s * 3 * n_i;
#pragma unroll
        for (int c = 0; c < 3; ++c) a[c] += __ldcg(p + (size_t)c * n_i + li);  // written by other CTAs: read at L2
    }
    if (e.mode == kEpiAccel) {
#pragma unroll
        for (int c = 0; c < 3; ++c) e.acc[(size_t)li * 3 + c] = a[c];
        return;
    }
    const int gi = i0 + li;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T x = StreamIO<T>::get(e.cur, gi, c);
        T v = e.vel[(size_t)li * 3 + c];
        v = mul_add_unfused(e.half_dt, a[c], v);  // closing kick, nbody.py:214
        if (e.flags & NB_STEP_SNAPSHOT) {          // get_state(), nbody.py:250-259
            if (e.sp) e.sp[(size_t)gi * 3 + c] = (double)x;
            if (e.sv) e.sv[(size_t)gi * 3 + c] = (double)v;
            if (e.sa) e.sa[(size_t)gi * 3 + c] = (double)a[c];
        }
        if (e.flags & NB_STEP_CONTINUE) {
            v = mul_add_unfused(e.half_dt, a[c], v);  // next step's opening kick, nbody.py:205
            x = mul_add_unfused(e.dt, v, x);          // drift, nbody.py:208
            if (e.mode == kEpiStep) StreamIO<T>::put(e.next, gi, c, x);
        }
        x_out[c] = x;
        e.vel[(size_t)li * 3 + c] = v;
        e.acc[(size_t)li * 3 + c] = a[c];
    }
}

// Store one 32-byte stream record into the next stream of every rank (two 16-byte stores per peer over NVLink).
template <typename V>
__device__ __forceinline__ void store_record_to_peers(const PeerTargets& peers, size_t rec16, V lo, V hi) {
    for (int p = 0; p < peers.n_ranks; ++p) {
        V* dst = reinterpret_cast<V*>(peers.next[p]) + rec16;
        dst[0] = lo;
        dst[1] = hi;
    }
}
__device__ __forceinline__ void peer_records(const Epilogue<double>& e, int gi, bool valid, const double x[3]) {
    if (valid && (e.flags & NB_STEP_CONTINUE)) {
        const double gm = e.cur[(size_t)gi * 4 + 3];
        store_record_to_peers<double2>(e.peers, (size_t)gi * 2, make_double2(x[0], x[1]), make_double2(x[2], gm));
    }
}
__device__ __forceinline__ void peer_records(const Epilogue<float>& e, int gi, bool valid, const float x[3]) {
    // a record holds the pair (even body, odd body): the even lane collects its neighbour's new position
    const float x1 = __shfl_down_sync(0xffffffffu, x[0], 1);
    const float y1 = __shfl_down_sync(0xffffffffu, x[1], 1);
    const float z1 = __shfl_down_sync(0xffffffffu, x[2], 1);
    NB_CHECK(!valid || gi >= 0);
    if (valid && !(gi & 1) && (e.flags & NB_STEP_CONTINUE)) {
        const float4 b = reinterpret_cast<const float4*>(e.cur)[(size_t)gi + 1];  // z0 z1 gm0 gm1 of this pair
        store_record_to_peers<float4>(e.peers, (size_t)gi, make_float4(x[0], x1, x[1], y1),
                                      make_float4(x[2], z1, b.z, b.w));
    }
}

// Called by every thread of a force CTA after its partials are stored.  The CTA that arrives last at its
// i-tile's counter owns the tile's finish: the force pass and the leapfrog are ONE launch.
template <typename T, int kP, int kBlock>
__device__ __forceinline__ void tile_epilogue(const Epilogue<T>& e, const T* partial, int i0, int n_i, int li0) {
    if (e.mode == kEpiNone) return;
    __shared__ int s_last;
    __threadfence();  // this thread's partials are visible device-wide before the arrival count
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(e.tile_counter + blockIdx.x, 1) == (int)gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
#ifdef NB_DEBUG_CHECKS
    NB_CHECK(*reinterpret_cast<volatile int*>(e.tile_counter + blockIdx.x) == (int)gridDim.y);  // every segment arrived once
    __syncthreads();
#endif
    if (threadIdx.x == 0) e.tile_counter[blockIdx.x] = 0;  // ready for the next launch
    __threadfence();
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int li = li0 + k * kBlock;
        const bool valid = li < n_i;
        T x[3] = {T(0), T(0), T(0)};  // lanes past the slab stand for padding bodies (position 0, G*m 0)
        if (valid) finish_body<T>(e, partial, i0, n_i, li, x);
        if (e.mode == kEpiStepPeer) peer_records(e, i0 + li, valid, x);
    }
    if (e.mode != kEpiStepPeer) return;
    // K2 fused with its collective.  This tile's records are on their way to every rank; when the LAST tile of
    // the launch gets here, one thread per peer publishes this rank's arrival word there and (NB_STEP_PEER_SYNC)
    // waits for that peer's word here, so the kernel boundary orders the next force pass after every rank's stores.
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(e.done_counter, 1) == (int)gridDim.x - 1;
        if (s_last) *e.done_counter = 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x < e.peers.n_ranks) {
        __threadfence_system();
        st_release_sys(e.peers.flags[threadIdx.x] + e.peers.my_rank, e.peers.seq);
        if (e.flags & NB_STEP_PEER_SYNC) {
            const uint32_t* mine = e.peers.flags[e.peers.my_rank] + threadIdx.x;
            const unsigned long long t0 = global_ns();
            while ((int32_t)(ld_acquire_sys(mine) - e.peers.seq) < 0) {
                __nanosleep(50);
                if (global_ns() - t0 > e.peers.timeout_ns) {  // that rank died: say so, do not hang the GPU
                    peer_lost(e.peers.error, NB_PEER_LOST_AFTER_STEP, (int)threadIdx.x);
                    break;
                }
            }
        }
    }
}

// __launch_bounds__(kBlock, 512 / kBlock): two 256-thread (or four 128-thread) CTAs per SM and up to 128
// registers; without the second argument ptxas settles for 76 registers and a 3% slower schedule.
template <int kP, int kBlock, bool kZeroEps>
__global__ void __launch_bounds__(kBlock, 512 / kBlock)
force_f32_kernel(const float* __restrict__ stream, int n_pad, int i0, int n_i, int seg_len, float eps2,
                 float* __restrict__ partial, const PeerWait wait, const Epilogue<float> epi) {
    __shared__ __align__(128) char ring[kStages * kTileBytes];
    __shared__ __align__(8) uint64_t bars[kStages];
    pdl_prologue();
    if (!peer_wait(wait)) return;

    const int seg = blockIdx.y;
    const int j0 = seg * seg_len;
    const int j1 = min(j0 + seg_len, n_pad);
    const int li0 = blockIdx.x * (kBlock * kP) + threadIdx.x;
    NB_CHECK(j0 < j1 && (j0 % kChunkBodies) == 0 && (j1 % kChunkBodies) == 0 && i0 >= 0 && i0 + n_i <= n_pad);

    float xi[kP], yi[kP], zi[kP];
    float2 ax[kP], ay[kP], az[kP];
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int gi = i0 + min(li0 + k * kBlock, n_i - 1);
        xi[k] = f32_stream_coord(stream, gi, 0);
        yi[k] = f32_stream_coord(stream, gi, 1);
        zi[k] = f32_stream_coord(stream, gi, 2);
        ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
    }
    // tiles that overlap this CTA's own bodies take the guarded copy of the loop (nb_tiles.cuh, f32_pairs)
    const int own_lo = i0 + blockIdx.x * (kBlock * kP), own_hi = min(own_lo + kBlock * kP, i0 + n_i);
    auto consume = [&](const char* tile, int bytes, int tile_index) {
        const float4* __restrict__ t = reinterpret_cast<const float4*>(tile);
        const int n_pairs = bytes >> 5;
        const int jt_lo = j0 + tile_index * (kTileBytes / 16), jt_hi = jt_lo + 2 * n_pairs;
        if (jt_lo < own_hi && own_lo < jt_hi) f32_pairs<kP, true>(t, n_pairs, xi, yi, zi, ax, ay, az, eps2);
        else f32_pairs<kP, false>(t, n_pairs, xi, yi, zi, ax, ay, az, eps2);
    };
    stream_tiles(reinterpret_cast<const char*>(stream) + (size_t)j0 * 16, (j1 - j0) * 16, ring, bars, consume);
    pdl_release();

    float* __restrict__ out = partial + (size_t)seg * 3 * n_i;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int li = li0 + k * kBlock;
        if (li < n_i) {
            out[li] = ax[k].x + ax[k].y;  // even-j lane + odd-j lane
            out[(size_t)n_i + li] = ay[k].x + ay[k].y;
            out[(size_t)2 * n_i + li] = az[k].x + az[k].y;
        }
    }
    tile_epilogue<float, kP, kBlock>(epi, partial, i0, n_i, li0);
}

// ------------------------------------------------------------------------------------------------
// float64 force kernel
// ------------------------------------------------------------------------------------------------
template <int kP, int kBlock, bool kZeroEps>
__global__ void __launch_bounds__(kBlock)
force_f64_kernel(const double* __restrict__ stream, int n_pad, int i0, int n_i, int seg_len, double eps2,
                 double* __restrict__ partial, const PeerWait wait, const Epilogue<double> epi) {
    __shared__ __align__(128) char ring[kStages * kTileBytes];
    __shared__ __align__(8) uint64_t bars[kStages];
    pdl_prologue();
    if (!peer_wait(wait)) return;

    const int seg = blockIdx.y;
    const int j0 = seg * seg_len;
    const int j1 = min(j0 + seg_len, n_pad);
    const int li0 = blockIdx.x * (kBlock * kP) + threadIdx.x;

    double xi[kP], yi[kP], zi[kP], ax[kP], ay[kP], az[kP];
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int gi = i0 + min(li0 + k * kBlock, n_i - 1);
        const double4 q = reinterpret_cast<const double4*>(stream)[gi];
        xi[k] = q.x;
        yi[k] = q.y;
        zi[k] = q.z;
        ax[k] = ay[k] = az[k] = 0.0;
    }

    auto consume = [&](const char* tile, int bytes, int) {
        const double2* __restrict__ t = reinterpret_cast<const double2*>(tile);
        f64_bodies<kP, kZeroEps>(t, bytes >> 5, xi, yi, zi, ax, ay, az, eps2);
    };
    stream_tiles(reinterpret_cast<const char*>(stream) + (size_t)j0 * 32, (j1 - j0) * 32, ring, bars, consume);
    pdl_release();

    double* __restrict__ out = partial + (size_t)seg * 3 * n_i;
#pragma unroll
    for (int k = 0; k < kP; ++k) {
        const int li = li0 + k * kBlock;
        if (li < n_i) {
            out[li] = ax[k];
            out[(size_t)n_i + li] = ay[k];
            out[(size_t)2 * n_i + li] = az[k];
        }
    }
    tile_epilogue<double, kP, kBlock>(epi, partial, i0, n_i, li0);
}

template <typename T>
__global__ void __launch_bounds__(256)
kick_drift_kernel(const T* __restrict__ stream_cur, T* __restrict__ stream_next, T* __restrict__ vel,
                  const T* __restrict__ acc, int i0, int n_i, T dt, T half_dt) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_i) return;
    const int gi = i0 + li;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T v = vel[(size_t)li * 3 + c];
        v = mul_add_unfused(half_dt, acc[(size_t)li * 3 + c], v);           // nbody.py:205
        const T x = mul_add_unfused(dt, v, StreamIO<T>::get(stream_cur, gi, c));  // nbody.py:208
        StreamIO<T>::put(stream_next, gi, c, x);
        vel[(size_t)li * 3 + c] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pack_kernel(const double* __restrict__ pos, const void* __restrict__ masses, int masses_are_f32, int n, int n_pad,
            T* __restrict__ stream) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    double x = 0.0, y = 0.0, z = 0.0, gm = 0.0;
    if (i < n) {
        x = pos[(size_t)i * 3 + 0];
        y = pos[(size_t)i * 3 + 1];
        z = pos[(size_t)i * 3 + 2];
        const double m = masses_are_f32 ? (double)static_cast<const float*>(masses)[i]
                                        : static_cast<const double*>(masses)[i];
        gm = kG * m;  // G * masses[j], nbody.py:57
        if (sizeof(T) == 8) gm *= kMassSlotF64;  // 1 unless the first-order f64 pair is built (nb_common.cuh)
    }
    StreamIO<T>::put(stream, i, 0, (T)x);
    StreamIO<T>::put(stream, i, 1, (T)y);
    StreamIO<T>::put(stream, i, 2, (T)z);
    StreamIO<T>::put(stream, i, 3, (T)gm);
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_kernel(const T* __restrict__ stream, int n, double* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) pos[(size_t)i * 3 + c] = (double)StreamIO<T>::get(stream, i, c);
}

// ------------------------------------------------------------------------------------------------
// host-side launch logic
// ------------------------------------------------------------------------------------------------
struct Slab {
    int n, n_pad, i0, n_i, seg_len, n_seg;
};

static int make_slab(int n, int i0, int n_i, Slab* out) {
    NB_REQUIRE(n > 0, "n must be positive (got %d)", n);
    NB_REQUIRE(i0 >= 0 && n_i > 0 && i0 + n_i <= n, "slab [%d, %d) outside system of %d bodies", i0, i0 + n_i, n);
    out->n = n;
    out->n_pad = nb_padded_bodies(n);
    out->i0 = i0;
    out->n_i = n_i;
    nb_segment_plan(n, &out->seg_len, &out->n_seg);
    return NB_OK;
}

// i-tile shapes.  Small slabs take the small tile so the grid still covers the SMs.
template <typename T>
struct Tile;
template <>
struct Tile<float> {
    static constexpr int kPBig = 4, kBlockBig = 256, kPSmall = 2, kBlockSmall = 128;
};
template <>
struct Tile<double> {
    static constexpr int kPBig = 2, kBlockBig = 256, kPSmall = 1, kBlockSmall = 128;
};

// Launch with programmatic stream serialization (see pdl_prologue): the kernel may be scheduled before its
// predecessor on the stream has finished; it orders itself with griddepcontrol.wait.
template <class Kernel, class... Args>
static void launch_pdl(Kernel kern, dim3 grid, int block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int kP, int kBlock, bool kZeroEps>
static void launch_force(const float* stream, const Slab& sl, float eps2, float* partial, const PeerWait& w,
                         const Epilogue<float>& e, cudaStream_t st) {
    dim3 grid(ceil_div(sl.n_i, kP * kBlock), sl.n_seg);
    launch_pdl(force_f32_kernel<kP, kBlock, false>, grid  /* one build: the guard is unconditional */, kBlock, st, stream, sl.n_pad, sl.i0, sl.n_i, sl.seg_len,
               eps2, partial, w, e);
}
template <int kP, int kBlock, bool kZeroEps>
static void launch_force(const double* stream, const Slab& sl, double eps2, double* partial, const PeerWait& w,
                         const Epilogue<double>& e, cudaStream_t st) {
    dim3 grid(ceil_div(sl.n_i, kP * kBlock), sl.n_seg);
    launch_pdl(force_f64_kernel<kP, kBlock, kZeroEps>, grid, kBlock, st, stream, sl.n_pad, sl.i0, sl.n_i, sl.seg_len,
               eps2, partial, w, e);
}

// Workspace layout: [i-tile arrival counters, sized by n alone][segment partials of this slab].  The header does
// not depend on the slab, so one workspace serves calls on different slabs of the same system.
static size_t counter_bytes(int n) { return ws_header_bytes(n); }
static int* tile_counters(void* ws) { return static_cast<int*>(ws); }
static int* done_counter(void* ws, int n) { return static_cast<int*>(ws) + ((size_t)n / 128 + 1); }
static int* error_word(void* ws, int n) { return static_cast<int*>(ws) + ((size_t)n / 128 + 2); }

// How long a kernel of the sharded mode waits for another rank before it declares it lost (NB_PEER_TIMEOUT_MS,
// default 10 s; read once).
static unsigned long long peer_timeout_ns() {
    static unsigned long long ns = 0;
    if (ns == 0) {
        const char* e = getenv("NB_PEER_TIMEOUT_MS");
        double ms = e ? atof(e) : 0.0;
        if (!(ms > 0.0)) ms = 10000.0;
        ns = (unsigned long long)(ms * 1e6);
    }
    return ns;
}
template <typename T>
static T* partials(void* ws, int n) { return reinterpret_cast<T*>(static_cast<char*>(ws) + counter_bytes(n)); }

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <typename T>
static int force_pass(const T* stream, const Slab& sl, double softening, T* partial, cudaStream_t st, Epilogue<T> e,
                      const PeerWait& w = PeerWait{nullptr, 0, 0, 0ull, nullptr}) {
    e.n_seg = sl.n_seg;
    const T eps2 = (T)(softening * softening);
    const bool zero = !(eps2 > T(0));
    using TL = Tile<T>;
    // big tile only when it still yields at least ~4 CTAs per SM
    const long ctas_big = (long)ceil_div(sl.n_i, TL::kPBig * TL::kBlockBig) * sl.n_seg;
    const bool big = ctas_big >= 4L * sm_count();
    if (big) {
        if (zero) launch_force<TL::kPBig, TL::kBlockBig, true>(stream, sl, eps2, partial, w, e, st);
        else launch_force<TL::kPBig, TL::kBlockBig, false>(stream, sl, eps2, partial, w, e, st);
    } else {
        if (zero) launch_force<TL::kPSmall, TL::kBlockSmall, true>(stream, sl, eps2, partial, w, e, st);
        else launch_force<TL::kPSmall, TL::kBlockSmall, false>(stream, sl, eps2, partial, w, e, st);
    }
    return check_launch("force kernel");
}

// Whole systems small enough for one wave of group CTAs take K2s (nb_group.cu: no cross-CTA reduction, the same bits);
// NB_NO_GROUP=1 keeps K2 (the tests compare the two).
static int whole_system_kp(const Slab& sl, int is_f64) {
    if (sl.i0 != 0 || sl.n_i != sl.n) return 0;
    const char* e = getenv("NB_NO_GROUP");
    if (e && e[0] == '1') return 0;
    return group_step_kp(sl.n, sl.n_seg, sm_count(), is_f64);
}

template <typename T>
static int accel_impl(const T* stream, int n, int i0, int n_i, double softening, T* acc, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
    Slab sl;
    if (int rc = make_slab(n, i0, n_i, &sl)) return rc;
    NB_REQUIRE(stream && acc && ws, "null pointer argument");
    NB_REQUIRE(ws_bytes >= nb_workspace_bytes(n, n_i, sizeof(T) == 8), "workspace too small: %zu < %zu", ws_bytes,
               nb_workspace_bytes(n, n_i, sizeof(T) == 8));
    if (const int kp = whole_system_kp(sl, sizeof(T) == 8))
        return group_step<T>(stream, nullptr, nullptr, acc, n, 0.0, softening, /*mode=*/1, 0, nullptr, nullptr, nullptr,
                             error_word(ws, n), kp, st);
    T* partial = partials<T>(ws, n);
    Epilogue<T> e{};
    e.mode = kEpiAccel;
    e.tile_counter = tile_counters(ws);
    e.acc = acc;
    return force_pass<T>(stream, sl, softening, partial, st, e);
}

template <typename T>
static int step_impl(const T* cur, T* next, T* vel, T* acc, int n, int i0, int n_i, double dt, double softening,
                     int flags, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes, cudaStream_t st) {
    Slab sl;
    if (int rc = make_slab(n, i0, n_i, &sl)) return rc;
    NB_REQUIRE(cur && vel && acc && ws, "null pointer argument");
    NB_REQUIRE(!(flags & NB_STEP_CONTINUE) || next, "NB_STEP_CONTINUE needs stream_next");
    NB_REQUIRE(ws_bytes >= nb_workspace_bytes(n, n_i, sizeof(T) == 8), "workspace too small: %zu < %zu", ws_bytes,
               nb_workspace_bytes(n, n_i, sizeof(T) == 8));
    if (const int kp = whole_system_kp(sl, sizeof(T) == 8))
        return group_step<T>(cur, next, vel, acc, n, dt, softening, /*mode=*/2, flags, sp, sv, sa, error_word(ws, n), kp,
                             st);
    T* partial = partials<T>(ws, n);
    const double half_dt = 0.5 * dt;  // "0.5 * self.dt" is evaluated first, nbody.py:205
    Epilogue<T> e{};
    e.mode = kEpiStep;
    e.tile_counter = tile_counters(ws);
    e.cur = cur; e.next = next; e.vel = vel; e.acc = acc;
    e.dt = (T)dt; e.half_dt = (T)half_dt; e.flags = flags;
    e.sp = sp; e.sv = sv; e.sa = sa;
    return force_pass<T>(cur, sl, softening, partial, st, e);
}

template <typename T>
static int step_peer_impl(const T* cur, void* const* next_peers, void* const* flag_peers, int n_ranks, int my_rank,
                          unsigned wait_seq, unsigned signal_seq, T* vel, T* acc, int n, int i0, int n_i, double dt,
                          double softening, int flags, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
    Slab sl;
    if (int rc = make_slab(n, i0, n_i, &sl)) return rc;
    NB_REQUIRE(cur && vel && acc && ws && next_peers && flag_peers, "null pointer argument");
    NB_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxPeers && my_rank >= 0 && my_rank < n_ranks,
               "need 1 <= n_ranks <= %d and 0 <= my_rank < n_ranks", kMaxPeers);
    NB_REQUIRE(i0 % kChunkBodies == 0, "slab start must be a multiple of %d bodies", kChunkBodies);
    NB_REQUIRE(ws_bytes >= nb_workspace_bytes(n, n_i, sizeof(T) == 8), "workspace too small: %zu < %zu", ws_bytes,
               nb_workspace_bytes(n, n_i, sizeof(T) == 8));
    PeerTargets tg;
    for (int p = 0; p < n_ranks; ++p) {
        NB_REQUIRE(next_peers[p] && flag_peers[p], "null peer pointer for rank %d", p);
        tg.next[p] = next_peers[p];
        tg.flags[p] = static_cast<uint32_t*>(flag_peers[p]);
    }
    tg.n_ranks = n_ranks; tg.my_rank = my_rank; tg.seq = signal_seq;
    tg.timeout_ns = peer_timeout_ns();
    tg.error = error_word(ws, n);
    PeerWait w{wait_seq ? tg.flags[my_rank] : nullptr, n_ranks, wait_seq, tg.timeout_ns, tg.error};
    T* partial = partials<T>(ws, n);
    const double half_dt = 0.5 * dt;
    Epilogue<T> e{};
    e.mode = kEpiStepPeer;
    e.tile_counter = tile_counters(ws);
    e.done_counter = done_counter(ws, n);
    e.peers = tg;
    e.cur = cur; e.vel = vel; e.acc = acc;
    e.dt = (T)dt; e.half_dt = (T)half_dt; e.flags = flags;
    e.sp = sp; e.sv = sv; e.sa = sa;
    return force_pass<T>(cur, sl, softening, partial, st, e, w);
}

template <typename T>
static int kick_drift_impl(const T* cur, T* next, T* vel, const T* acc, int n, int i0, int n_i, double dt,
                           cudaStream_t st) {
    Slab sl;
    if (int rc = make_slab(n, i0, n_i, &sl)) return rc;
    NB_REQUIRE(cur && next && vel && acc, "null pointer argument");
    const double half_dt = 0.5 * dt;
    kick_drift_kernel<T><<<ceil_div(n_i, 256), 256, 0, st>>>(cur, next, vel, acc, i0, n_i, (T)dt, (T)half_dt);
    return check_launch("kick_drift kernel");
}

template <typename T>
__global__ void __launch_bounds__(256)
snapshot0_kernel(const T* __restrict__ stream, const T* __restrict__ vel, const T* __restrict__ acc, int n,
                 double* __restrict__ sp, double* __restrict__ sv, double* __restrict__ sa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (sp) sp[(size_t)i * 3 + c] = (double)StreamIO<T>::get(stream, i, c);
        if (sv) sv[(size_t)i * 3 + c] = (double)vel[(size_t)i * 3 + c];
        if (sa) sa[(size_t)i * 3 + c] = (double)acc[(size_t)i * 3 + c];
    }
}

template <typename T>
static int run_impl(T* sa_, T* sb_, T* vel, T* acc, int n, double dt, double softening, int n_steps, int save_interval,
                    double* sp, double* sv, double* sa, void* ws, size_t ws_bytes, int* final_in_a, cudaStream_t st) {
    NB_REQUIRE(n_steps >= 0 && save_interval >= 1, "n_steps >= 0 and save_interval >= 1 required");
    NB_REQUIRE(sa_ && sb_ && vel && acc, "null pointer argument");
    const size_t row = (size_t)n * 3;
    if (sp || sv || sa) {  // states.append(get_state()) before the loop, nbody.py:235
        snapshot0_kernel<T><<<ceil_div(n, 256), 256, 0, st>>>(sa_, vel, acc, n, sp, sv, sa);
        if (int rc = check_launch("snapshot0 kernel")) return rc;
    }
    T* cur = sa_;
    T* next = sb_;
    if (n_steps > 0) {
        if (int rc = kick_drift_impl<T>(cur, next, vel, acc, n, 0, n, dt, st)) return rc;
        T* t = cur; cur = next; next = t;
    }
    // Opt-in (environment NB_PERSIST=1): every step of the run in ONE cooperative launch (nb_persist.cu) -- the same
    // bits as the per-step launches below.  Measured (profiles/r02_midn_persist.log): no faster than they are -- below
    // N ~ 8k a step is the same chain of L2 round trips either way, above it the per-warp rings feed the pipes worse
    // than K2's CTA-wide ring -- so the per-step path stays the default.
    const char* persist_env = getenv("NB_PERSIST");
    if (n_steps >= 2 && n <= nb_persist_max_bodies() && ws && persist_env && persist_env[0] == '1' &&
        ws_bytes >= nb_workspace_bytes(n, n, sizeof(T) == 8)) {
        int n_seg = 1;
        nb_segment_plan(n, nullptr, &n_seg);
        const size_t part_bytes = ((size_t)n_seg * 3 * n * sizeof(T) + 255) / 256 * 256;
        char* region = static_cast<char*>(ws) + counter_bytes(n) + part_bytes;
        unsigned* barrier = reinterpret_cast<unsigned*>(region);
        int* group_counter = reinterpret_cast<int*>(region) + 8;
        if (int rc = persist_run<T>(cur, next, vel, acc, n, dt, softening, n_steps, save_interval, sp, sv, sa,
                                    partials<T>(ws, n), group_counter, barrier, error_word(ws, n), st))
            return rc;
        // the streams alternate once per step but the last: where x_n is
        if ((n_steps - 1) % 2 == 1) { T* t = cur; cur = next; next = t; }
        if (final_in_a) *final_in_a = (cur == sa_) ? 1 : 0;
        return NB_OK;
    }
    size_t snap = 1;
    for (int k = 1; k <= n_steps; ++k) {
        int flags = 0;
        if (k < n_steps) flags |= NB_STEP_CONTINUE;
        const bool save = (k % save_interval) == 0;  // nbody.py:240
        if (save) flags |= NB_STEP_SNAPSHOT;
        if (int rc = step_impl<T>(cur, next, vel, acc, n, 0, n, dt, softening, flags,
                                  (save && sp) ? sp + snap * row : nullptr, (save && sv) ? sv + snap * row : nullptr,
                                  (save && sa) ? sa + snap * row : nullptr, ws, ws_bytes, st))
            return rc;
        if (save) ++snap;
        if (k < n_steps) { T* t = cur; cur = next; next = t; }
    }
    if (final_in_a) *final_in_a = (cur == sa_) ? 1 : 0;
    return NB_OK;
}

}  // namespace nb

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int nb_pack_f64(const double* pos, const void* masses, int masses_are_f32, int n, double* stream, nb_stream_t s) {
    NB_REQUIRE(pos && masses && stream && n > 0, "nb_pack_f64: bad argument");
    const int n_pad = nb_padded_bodies(n);
    nb::pack_kernel<double><<<nb::ceil_div(n_pad, 256), 256, 0, (cudaStream_t)s>>>(pos, masses, masses_are_f32, n,
                                                                                    n_pad, stream);
    return nb::check_launch("pack kernel");
}
int nb_pack_f32(const double* pos, const void* masses, int masses_are_f32, int n, float* stream, nb_stream_t s) {
    NB_REQUIRE(pos && masses && stream && n > 0, "nb_pack_f32: bad argument");
    const int n_pad = nb_padded_bodies(n);
    nb::pack_kernel<float><<<nb::ceil_div(n_pad, 256), 256, 0, (cudaStream_t)s>>>(pos, masses, masses_are_f32, n,
                                                                                   n_pad, stream);
    return nb::check_launch("pack kernel");
}
int nb_unpack_f64(const double* stream, int n, double* pos, nb_stream_t s) {
    NB_REQUIRE(pos && stream && n > 0, "nb_unpack_f64: bad argument");
    nb::unpack_kernel<double><<<nb::ceil_div(n, 256), 256, 0, (cudaStream_t)s>>>(stream, n, pos);
    return nb::check_launch("unpack kernel");
}
int nb_unpack_f32(const float* stream, int n, double* pos, nb_stream_t s) {
    NB_REQUIRE(pos && stream && n > 0, "nb_unpack_f32: bad argument");
    nb::unpack_kernel<float><<<nb::ceil_div(n, 256), 256, 0, (cudaStream_t)s>>>(stream, n, pos);
    return nb::check_launch("unpack kernel");
}

int nb_accel_f64(const double* stream, int n, int i0, int n_i, double softening, double* acc, void* ws,
                 size_t ws_bytes, nb_stream_t s) {
    return nb::accel_impl<double>(stream, n, i0, n_i, softening, acc, ws, ws_bytes, (cudaStream_t)s);
}
int nb_accel_f32(const float* stream, int n, int i0, int n_i, double softening, float* acc, void* ws, size_t ws_bytes,
                 nb_stream_t s) {
    return nb::accel_impl<float>(stream, n, i0, n_i, softening, acc, ws, ws_bytes, (cudaStream_t)s);
}

int nb_kick_drift_f64(const double* cur, double* next, double* vel, const double* acc, int n, int i0, int n_i,
                      double dt, nb_stream_t s) {
    return nb::kick_drift_impl<double>(cur, next, vel, acc, n, i0, n_i, dt, (cudaStream_t)s);
}
int nb_kick_drift_f32(const float* cur, float* next, float* vel, const float* acc, int n, int i0, int n_i, double dt,
                      nb_stream_t s) {
    return nb::kick_drift_impl<float>(cur, next, vel, acc, n, i0, n_i, dt, (cudaStream_t)s);
}

int nb_step_f64(const double* cur, double* next, double* vel, double* acc, int n, int i0, int n_i, double dt,
                double softening, int flags, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
                nb_stream_t s) {
    return nb::step_impl<double>(cur, next, vel, acc, n, i0, n_i, dt, softening, flags, sp, sv, sa, ws, ws_bytes,
                                 (cudaStream_t)s);
}
int nb_step_f32(const float* cur, float* next, float* vel, float* acc, int n, int i0, int n_i, double dt,
                double softening, int flags, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
                nb_stream_t s) {
    return nb::step_impl<float>(cur, next, vel, acc, n, i0, n_i, dt, softening, flags, sp, sv, sa, ws, ws_bytes,
                                (cudaStream_t)s);
}

int nb_step_peer_f64(const double* cur, void* const* next_peers, void* const* flag_peers, int n_ranks, int my_rank,
                     unsigned wait_seq, unsigned signal_seq, double* vel, double* acc, int n, int i0, int n_i, double dt,
                     double softening, int flags, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
                     nb_stream_t s) {
    return nb::step_peer_impl<double>(cur, next_peers, flag_peers, n_ranks, my_rank, wait_seq, signal_seq, vel, acc, n,
                                      i0, n_i, dt, softening, flags, sp, sv, sa, ws, ws_bytes, (cudaStream_t)s);
}
int nb_step_peer_f32(const float* cur, void* const* next_peers, void* const* flag_peers, int n_ranks, int my_rank,
                     unsigned wait_seq, unsigned signal_seq, float* vel, float* acc, int n, int i0, int n_i, double dt,
                     double softening, int flags, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
                     nb_stream_t s) {
    return nb::step_peer_impl<float>(cur, next_peers, flag_peers, n_ranks, my_rank, wait_seq, signal_seq, vel, acc, n,
                                     i0, n_i, dt, softening, flags, sp, sv, sa, ws, ws_bytes, (cudaStream_t)s);
}

int nb_step_status(const void* workspace, int n, nb_stream_t s) {
    NB_REQUIRE(workspace && n > 0, "nb_step_status: bad argument");
    int word = 0;
    NB_CUDA_OK(cudaMemcpyAsync(&word, nb::error_word(const_cast<void*>(workspace), n), sizeof(int),
                               cudaMemcpyDeviceToHost, (cudaStream_t)s));
    NB_CUDA_OK(cudaStreamSynchronize((cudaStream_t)s));
    if (word == 0) return NB_OK;
    const int phase = word & 0xff, rank = word >> 8;
    if (phase == NB_PERSIST_STALLED) {
        nb::set_error("one-launch run: a grid barrier or a tile copy never completed; the state of this system is no "
                      "longer valid");
        return NB_ERR_CUDA;
    }
    nb::set_error("sharded step: rank %d never arrived (%s); the state of this system is no longer valid", rank,
                  phase == NB_PEER_LOST_BEFORE_FORCE ? "its positions of the previous step were not published in time"
                                                     : "it did not finish the step in time");
    return NB_ERR_PEER;
}

int nb_run_f64(double* stream_a, double* stream_b, double* vel, double* acc, int n, double dt, double softening,
               int n_steps, int save_interval, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
               int* final_in_a, nb_stream_t s) {
    return nb::run_impl<double>(stream_a, stream_b, vel, acc, n, dt, softening, n_steps, save_interval, sp, sv, sa, ws,
                                ws_bytes, final_in_a, (cudaStream_t)s);
}
int nb_run_f32(float* stream_a, float* stream_b, float* vel, float* acc, int n, double dt, double softening,
               int n_steps, int save_interval, double* sp, double* sv, double* sa, void* ws, size_t ws_bytes,
               int* final_in_a, nb_stream_t s) {
    return nb::run_impl<float>(stream_a, stream_b, vel, acc, n, dt, softening, n_steps, save_interval, sp, sv, sa, ws,
                               ws_bytes, final_in_a, (cudaStream_t)s);
}

}  // extern "C"
