/*
 * nbody_b200.h -- C ABI of libnbody_b200.so: direct-sum softened gravity + kick-drift-kick
 * leapfrog for NVIDIA B200 (sm_100a).
 *
 * The reference (Sanshrey712/nbody-gnn-hpc) has no FFI of its own: the seam is the Python
 * module `hpc.nbody` (reference src/hpc/nbody.py).  Each entry point below names the reference
 * code it replaces; `nbody-gnn-hpc_b200/hpc/nbody.py` is the Python mirror that binds them with
 * ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; `nb_stream_t` is a cudaStream_t passed as void*.
 *   - `nb_*`  : DEVICE pointers, stream-ordered, never allocate, never synchronise.
 *   - `nbh_*` : HOST pointers; allocate device scratch, copy in, run, copy out, synchronise.
 *   - return NB_OK (0) or an NB_ERR_* code; nb_last_error() gives the message for the calling
 *     thread.  Nothing throws; nothing falls back to a CPU path.
 *   - "API layout" means the reference's arrays: positions/velocities/accelerations (N,3)
 *     row-major float64, masses (N,) float64 or float32 (generate_data.py:109 assigns float32).
 *   - "stream layout" is the device-resident position+mass array the force kernels read:
 *       f64: n_pad x {x, y, z, G*m}                       (32 B per body)
 *       f32: n_pad/2 pairs x {x0,x1,y0,y1, z0,z1,G*m0,G*m1} (32 B per pair; feeds f32x2 math)
 *     n_pad = nb_padded_bodies(n); padding bodies have G*m = 0.
 */
#ifndef NBODY_B200_H_
#define NBODY_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 1
#define NB_G 6.67430e-11      /* reference src/hpc/nbody.py:18 */
#define NB_SOFTENING 1e-9     /* reference src/hpc/nbody.py:19 */
#define NB_CHUNK_BODIES 32    /* granularity of padding, j-segments and bulk copies */

#define NB_OK 0
#define NB_ERR_INVALID 1      /* bad argument (message says which) */
#define NB_ERR_CUDA 2         /* a CUDA runtime call or launch failed */
#define NB_ERR_PEER 3         /* sharded mode: another rank never arrived (nb_step_status) */

/* phases recorded in the workspace's error word by nb_step_peer_* (low byte; the lost rank is above it) */
#define NB_PEER_LOST_BEFORE_FORCE 1   /* wait_seq never reached: the force pass did not run */
#define NB_PEER_LOST_AFTER_STEP 2     /* NB_STEP_PEER_SYNC: signal_seq of that rank never seen */
#define NB_PERSIST_STALLED 3          /* nb_run_* one-launch kernel: a grid barrier never completed */

/* nb_step_* flags */
#define NB_STEP_CONTINUE 1    /* after the closing kick, also do the next step's opening kick + drift */
#define NB_STEP_SNAPSHOT 2    /* write (x, v, a) of the synchronised state to the snapshot rows */
#define NB_STEP_PEER_SYNC 4   /* nb_step_peer_*: do not finish before every rank has published signal_seq here */

typedef void* nb_stream_t;

/* ---- library / device ------------------------------------------------------------------- */
int nb_abi_version(void);
const char* nb_last_error(void);
/* sm_count, compute capability and max SM clock (kHz) of `device`. */
int nb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz);

/* Measured FMA-pipe peak of the current device in TFLOP/s (2 flops per FMA): mode 0 = FFMA,
 * 1 = packed FFMA2 (f32x2), 2 = DFMA.  scratch: one float on the device.  Synchronises. */
int nb_probe_fma_peak(int mode, double* tflops, float* scratch, nb_stream_t s);

/* Test aid: n_ctas thread blocks that each hold smem_bytes of shared memory and spin for `milliseconds` (<= 5000)
 * on stream s -- "another tenant's kernel" for the tests of the launches that need all their blocks resident. */
int nb_probe_occupy(int n_ctas, size_t smem_bytes, double milliseconds, nb_stream_t s);

/* ---- planning: pure host arithmetic, callable without a GPU ------------------------------ */
/* Padded system length (multiple of NB_CHUNK_BODIES, at least one chunk). */
int nb_padded_bodies(int n);
/* j-segmentation of a system of n bodies: every i sums segment partials in ascending segment
 * order, so the result depends on n only -- not on the slab [i0, i0+n_i) or the rank count. */
int nb_segment_plan(int n, int* seg_len, int* n_seg);
/* Bytes of device scratch nb_accel_* / nb_step_* need for a slab of n_i bodies in a system of n.  The scratch must be
 * ZERO before the first call that uses it (it ends with one arrival counter per i-tile); every call leaves the counters
 * zero, so one cudaMemset at allocation time is enough. */
size_t nb_workspace_bytes(int n, int n_i, int is_f64);

/* ---- layout --------------------------------------------------------------------------------
 * API layout -> stream layout.  Replaces nothing in the reference (it has no device state);
 * G*m_j is formed here once, in float64 from the stored (possibly float32) mass, the product
 * the reference forms per pair at src/hpc/nbody.py:57.  Writes all n_pad entries. */
int nb_pack_f64(const double* pos, const void* masses, int masses_are_f32, int n, double* stream, nb_stream_t s);
int nb_pack_f32(const double* pos, const void* masses, int masses_are_f32, int n, float* stream, nb_stream_t s);
/* stream layout -> API-layout float64 positions (n,3). */
int nb_unpack_f64(const double* stream, int n, double* pos, nb_stream_t s);
int nb_unpack_f32(const float* stream, int n, double* pos, nb_stream_t s);

/* ---- K1: accelerations --------------------------------------------------------------------
 * Replaces compute_accelerations_direct, reference src/hpc/nbody.py:22-66, for the rows
 * [i0, i0+n_i) of a system of n bodies held in stream layout.  acc is (n_i,3) in the kernel's
 * dtype.  softening == 0 selects the guarded variant (i == j and exact overlaps contribute 0). */
int nb_accel_f64(const double* stream, int n, int i0, int n_i, double softening, double* acc,
                 void* workspace, size_t workspace_bytes, nb_stream_t s);
int nb_accel_f32(const float* stream, int n, int i0, int n_i, double softening, float* acc,
                 void* workspace, size_t workspace_bytes, nb_stream_t s);

/* ---- K2: fused leapfrog ---------------------------------------------------------------------
 * Replaces NBodySimulator.step, reference src/hpc/nbody.py:202-218, and the get_state copies
 * of run(), :235,:240-241,:250-259.
 *
 * nb_kick_drift_*: opening half of a step for the slab: v += (dt/2)*a ; x += dt*v, new positions
 *   written to rows [i0, i0+n_i) of stream_next (:205,:208).
 * nb_step_*: ONE launch: the force pass on stream_cur (positions x_k of all n bodies); the CTA that completes an
 *   i-tile last adds the segment partials in order and applies the closing
 *   kick v_k = v + (dt/2)*a_k (:211,:214), the optional snapshot of (x_k, v_k, a_k) into rows
 *   [i0, i0+n_i) of snap_pos/snap_vel/snap_acc (API layout float64, may be NULL), and -- with
 *   NB_STEP_CONTINUE -- the next step's opening kick and drift into stream_next.
 * vel and acc are (n_i,3) in the kernel's dtype; vel holds v_{k-1/2} on entry and v_k (or
 * v_{k+1/2} with NB_STEP_CONTINUE) on return; acc receives a_k.
 * The f64 integrator rounds every product and sum separately, as NumPy does. */
int nb_kick_drift_f64(const double* stream_cur, double* stream_next, double* vel, const double* acc,
                      int n, int i0, int n_i, double dt, nb_stream_t s);
int nb_kick_drift_f32(const float* stream_cur, float* stream_next, float* vel, const float* acc,
                      int n, int i0, int n_i, double dt, nb_stream_t s);
int nb_step_f64(const double* stream_cur, double* stream_next, double* vel, double* acc,
                int n, int i0, int n_i, double dt, double softening, int flags,
                double* snap_pos, double* snap_vel, double* snap_acc,
                void* workspace, size_t workspace_bytes, nb_stream_t s);
int nb_step_f32(const float* stream_cur, float* stream_next, float* vel, float* acc,
                int n, int i0, int n_i, double dt, double softening, int flags,
                double* snap_pos, double* snap_vel, double* snap_acc,
                void* workspace, size_t workspace_bytes, nb_stream_t s);

/* nb_step_* fused with its collective, for one system sharded by i-slab over the GPUs of a box (no
 * counterpart in the reference, which is single-host).  ONE launch, same arithmetic as nb_step_*: the i-tile
 * epilogue stores the slab's drifted stream records directly into the next-stream buffer of EVERY rank over NVLink
 * (peer pointers from symmetric / IPC-mapped memory) while other tiles are still computing, and ordering between
 * ranks is carried by arrival words that the launch's last tile exchanges, instead of a collective call:
 *   next_peers[r]   rank r's next-stream buffer (device pointer valid on this GPU), r = 0..n_ranks-1, own included
 *   flag_peers[r]   rank r's flag array: >= 16 uint32, zero before first use; word q = last sequence number rank q
 *                   published to rank r
 *   wait_seq        the force pass reads stream_cur only after every rank has published >= wait_seq here (0: no wait)
 *   signal_seq      published to every rank once this rank's whole slab has been stored (> 0, increasing per step)
 *   NB_STEP_PEER_SYNC in flags: the call additionally waits (on the device, in its last thread block) until every
 *                   rank has published signal_seq here, so the next launch needs no wait_seq
 * i0 must be a multiple of NB_CHUNK_BODIES.  Host arrays of n_ranks pointers; n_ranks <= 16.
 * A rank that does not arrive within NB_PEER_TIMEOUT_MS (environment, default 10000) is declared lost: the kernel
 * records it in the workspace's error word and stops -- this launch and every later one on the workspace leave
 * without computing -- and nb_step_status() reports it.  A lost peer is never a step from stale positions. */
int nb_step_peer_f64(const double* stream_cur, void* const* next_peers, void* const* flag_peers, int n_ranks,
                     int my_rank, unsigned wait_seq, unsigned signal_seq, double* vel, double* acc,
                     int n, int i0, int n_i, double dt, double softening, int flags,
                     double* snap_pos, double* snap_vel, double* snap_acc,
                     void* workspace, size_t workspace_bytes, nb_stream_t s);
int nb_step_peer_f32(const float* stream_cur, void* const* next_peers, void* const* flag_peers, int n_ranks,
                     int my_rank, unsigned wait_seq, unsigned signal_seq, float* vel, float* acc,
                     int n, int i0, int n_i, double dt, double softening, int flags,
                     double* snap_pos, double* snap_vel, double* snap_acc,
                     void* workspace, size_t workspace_bytes, nb_stream_t s);

/* Synchronises stream s and reports whether any nb_step_peer_* launch on this workspace lost a peer:
 * NB_OK, or NB_ERR_PEER with the rank and phase in nb_last_error().  Sticky until the workspace is zeroed. */
int nb_step_status(const void* workspace, int n, nb_stream_t s);

/* Replaces the loop of NBodySimulator.run, reference src/hpc/nbody.py:237-241, on one GPU:
 * enqueues nb_kick_drift + n_steps x nb_step on stream s.  stream_a holds x_0 on entry; the two
 * stream buffers alternate and *final_in_a reports which one holds x_{n_steps}.  vel/acc hold
 * (v_0, a_0) on entry and (v_n, a_n) on return.  Snapshot s (s = 1..n_steps/save_interval) of the
 * state after step s*save_interval goes to row block s of snap_* ((n_snap, n, 3) float64, n_snap =
 * 1 + n_steps/save_interval); row block 0 (the initial state) is written here too. */
/* Whole systems (i0 == 0, n_i == n) of at most 64 bodies per SM run nb_accel_* / nb_step_* / nb_run_* through K2s
 * (csrc/nb_group.cu: one thread block per group of bodies, the segment partials reduced in shared memory) -- the same
 * bits as the (i-tile x segment) kernels, which environment NB_NO_GROUP=1 selects for them too.
 * Opt-in, environment NB_PERSIST=1: systems of at most nb_persist_max_bodies() bodies run ALL their steps in one
 * cooperative launch (K2p, csrc/nb_persist.cu: warp-tasks on a persistent grid, one grid barrier per step) --
 * bit-identical to the per-step launches and, as measured, no faster; nb_step_status(workspace) reports
 * a stalled launch. */
int nb_persist_max_bodies(void);
int nb_run_f64(double* stream_a, double* stream_b, double* vel, double* acc, int n, double dt,
               double softening, int n_steps, int save_interval,
               double* snap_pos, double* snap_vel, double* snap_acc,
               void* workspace, size_t workspace_bytes, int* final_in_a, nb_stream_t s);
int nb_run_f32(float* stream_a, float* stream_b, float* vel, float* acc, int n, double dt,
               double softening, int n_steps, int save_interval,
               double* snap_pos, double* snap_vel, double* snap_acc,
               void* workspace, size_t workspace_bytes, int* final_in_a, nb_stream_t s);

/* ---- K3: ensemble of independent small systems ------------------------------------------------
 * Replaces the mp.Pool of generate_single_simulation workers, reference
 * scripts/generate_data.py:32-58,142-149 (and the loop of scripts/evaluate.py:81-93): B
 * independent systems of N bodies advanced n_steps inside ONE launch, positions resident in
 * shared memory, snapshots streamed to HBM.
 *   x, v, a     (B,N,3) float64 API layout, in/out: state on entry, state after n_steps on return.
 *   masses      (N) shared by all systems (mass_stride 0) or (B,N) (mass_stride N); f64 or f32.
 *   compute_a0  1: evaluate a from x first (generate_data.py:47); 0: trust a as given.
 *   out_*       (B, n_snap_total, N, 3) float64 or NULL.  Snapshots of this call occupy rows
 *               snap_offset .. ; row snap_offset is the entry state when write_initial != 0,
 *               then one row per save_interval steps.
 * N is limited by shared memory (nb_ensemble_max_bodies()).  f32 variant computes forces and
 * integrates in float32 and converts on the way in and out. */
int nb_ensemble_max_bodies(void);
/* Device scratch: one hand-over flag per system, for systems whose steps are shared by two neighbouring workers
 * of the interval schedule (zeroed by the call itself). */
size_t nb_ensemble_workspace_bytes(int B);
/* The interval schedule of the ensemble kernel, as host arithmetic (for tests and tooling): the pieces worker w of
 * n_workers (<= B) runs, in order; out[5*i..5*i+4] = {system, first step, last step, waits for the previous worker's
 * flag, parks its state for the next worker}; *n_pieces = their number (only the first max_pieces are written). */
int nb_ensemble_worker_plan(int B, int n_steps, int n_workers, int w, int* out, int max_pieces, int* n_pieces);
int nb_ensemble_f64(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride,
                    int B, int N, double dt, double softening, int n_steps, int save_interval,
                    int compute_a0, int write_initial,
                    double* out_x, double* out_v, double* out_a, int n_snap_total, int snap_offset,
                    void* workspace, size_t workspace_bytes, nb_stream_t s);
int nb_ensemble_f32(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride,
                    int B, int N, double dt, double softening, int n_steps, int save_interval,
                    int compute_a0, int write_initial,
                    double* out_x, double* out_v, double* out_a, int n_snap_total, int snap_offset,
                    void* workspace, size_t workspace_bytes, nb_stream_t s);

/* Ensembles of systems too large for one CTA's shared memory (N > nb_ensemble_max_bodies(), up to
 * nb_batched_max_bodies()): the same per-simulation work of generate_data.py:32-58 -- which the unchanged script asks
 * for with --particles in the thousands -- B systems side by side in every launch (K2s, csrc/nb_group.cu, grid =
 * groups x B), one launch per leapfrog step for the whole ensemble instead of one run per system.
 *   streams     (B, nb_stream_elems(n)) stream layout, one system after the other (nb_pack_* per system)
 *   vel, acc    (B, n, 3) in the kernel dtype; snap_*  (B, 1 + n_steps/save_interval, n, 3) float64 or all NULL
 * nb_accel_batched_*: a_0 of every system.  nb_run_batched_*: the run loop, as nb_run_* for one system. */
int nb_batched_max_bodies(void);
int nb_accel_batched_f64(const double* streams, int B, int n, double softening, double* acc, nb_stream_t s);
int nb_accel_batched_f32(const float* streams, int B, int n, double softening, float* acc, nb_stream_t s);
int nb_run_batched_f64(double* stream_a, double* stream_b, double* vel, double* acc, int B, int n, double dt,
                       double softening, int n_steps, int save_interval,
                       double* snap_pos, double* snap_vel, double* snap_acc, int* final_in_a, nb_stream_t s);
int nb_run_batched_f32(float* stream_a, float* stream_b, float* vel, float* acc, int B, int n, double dt,
                       double softening, int n_steps, int save_interval,
                       double* snap_pos, double* snap_vel, double* snap_acc, int* final_in_a, nb_stream_t s);

/* Strided device -> host copy of snapshot rows (height rows of width bytes; pitches in bytes), so the
 * rows of one step-chunk of every system can drain to pinned host memory while the next chunk runs. */
int nb_copy_rows_d2h_async(void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t width,
                           size_t height, nb_stream_t s);

/* ---- K4: energy ---------------------------------------------------------------------------------
 * Replaces compute_total_energy, reference src/hpc/nbody.py:101-130 (and the per-step energy of
 * src/utils/metrics.py:62-109): kinetic and potential energy contributed by rows [i0, i0+n_i) --
 * K = sum 1/2 m v^2, U = -1/2 sum_i m_i sum_{j != i} G m_j / sqrt(r^2 + eps^2) -- accumulated in
 * float64.  out_ku is 2 doubles on the device (K, U); ranks add their slabs.  API-layout inputs. */
size_t nb_energy_workspace_bytes(int n, int n_i);
int nb_energy_f64(const double* pos, const double* vel, const void* masses, int masses_are_f32,
                  int n, int i0, int n_i, double softening, double* out_ku,
                  void* workspace, size_t workspace_bytes, nb_stream_t s);

/* K4b: energy and momentum of EVERY snapshot of a stack of trajectories.  Replaces compute_energy_error, reference
 * src/utils/metrics.py:62-109 (a Python loop over the stored steps with an N x N x 3 temporary per step), and
 * compute_momentum_error, :112-137, for B trajectories x S snapshots in one launch:
 *   pos, vel   (B, S, N, 3) float64 device stacks (what nb_ensemble_* writes); pos may be NULL: K and p only, U = 0
 *   masses     (N) shared (mass_stride 0) or (B, N) (mass_stride N); f64 or f32 (f32: m_i*m_j is rounded to float32,
 *              as np.outer of the reference's float32 mass array is, metrics.py:82)
 *   G          the constant of metrics.py:65 (NB_G by default there)
 *   out        (B, S, 5) float64 device: K, U, px, py, pz of every snapshot
 * N <= nb_snapshot_energy_max_bodies() (a snapshot lives in shared memory); larger systems: nb_energy_f64 per state. */
int nb_snapshot_energy_max_bodies(void);
int nb_snapshot_energy_f64(const double* pos, const double* vel, const void* masses, int masses_are_f32,
                           int mass_stride, int B, int S, int N, double G, double softening, double* out,
                           nb_stream_t s);

/* ---- K5: sliding-window training samples --------------------------------------------------------
 * Replaces the sample loop of create_training_dataset, reference src/hpc/checkpoint.py:362-384
 * (and the sample count of :333): for every trajectory b of an ensemble whose float64 snapshot stacks
 * pos, vel (B, rows, N, 3) are in device memory, and every start i in range(0, n_states - L, stride),
 *   inputs [b*S + s] = concat(pos[b, i:i+L], vel[b, i:i+L], axis=-1) as float32   (L, N, 6)
 *   targets[b*S + s] = concat(pos[b, i+L],   vel[b, i+L],   axis=-1) as float32   (N, 6)
 * with S = nb_window_count(n_states, L, stride) samples per trajectory; n_states <= rows is the
 * trajectory's 'n_steps' entry (the number of stored states, generate_data.py:57). */
int nb_window_count(int n_states, int sequence_length, int stride);
int nb_window_gather_f32(const double* pos, const double* vel, int B, int rows, int N, int n_states,
                         int sequence_length, int stride, float* inputs, float* targets, nb_stream_t s);

/* ---- host-buffer entry points ---------------------------------------------------------------------
 * The same operations for callers that hold host arrays and no CUDA state (what a cgo/JNI/ctypes
 * binder of a host-language port would call).  Synchronous. */
int nbh_accel_direct(const double* pos, const void* masses, int masses_are_f32, int n, double softening,
                     int use_f32, double* acc);
int nbh_run(double* pos, double* vel, double* acc, const void* masses, int masses_are_f32, int n,
            double dt, double softening, int n_steps, int save_interval, int use_f32,
            double* snap_pos, double* snap_vel, double* snap_acc);
int nbh_ensemble_run(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride,
                     int B, int N, double dt, double softening, int n_steps, int save_interval, int use_f32,
                     double* out_x, double* out_v, double* out_a);
int nbh_total_energy(const double* pos, const double* vel, const void* masses, int masses_are_f32, int n,
                     double softening, double* out_kut /* K, U, K+U */);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H_ */
