// nb_host.cu -- host-buffer entry points (nbh_*): the same operations for callers that own host
// arrays only.  Each call allocates its device scratch, copies in, runs the nb_* device entry
// points on one stream, copies out and synchronises.  No CPU arithmetic happens here.
#include <vector>

#include "nb_common.cuh"

namespace nb {

// Frees every allocation of a call on scope exit, success or not.
struct DeviceArena {
    std::vector<void*> ptrs;
    cudaStream_t stream = nullptr;
    ~DeviceArena() {
        for (void* p : ptrs) cudaFree(p);
        if (stream) cudaStreamDestroy(stream);
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return NB_OK;
    }
};

#define NB_TRY(expr)                 \
    do {                             \
        if (int _rc = (expr)) return _rc; \
    } while (0)

template <typename T>
static int upload_as(T* dst, const double* src, size_t count, double* staging, cudaStream_t st);

template <>
int upload_as<double>(double* dst, const double* src, size_t count, double*, cudaStream_t st) {
    NB_CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, st));
    return NB_OK;
}

__global__ void narrow_kernel(const double* __restrict__ in, float* __restrict__ out, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (float)in[i];
}
__global__ void widen_kernel(const float* __restrict__ in, double* __restrict__ out, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (double)in[i];
}

template <>
int upload_as<float>(float* dst, const double* src, size_t count, double* staging, cudaStream_t st) {
    NB_CUDA_OK(cudaMemcpyAsync(staging, src, count * sizeof(double), cudaMemcpyHostToDevice, st));
    narrow_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(staging, dst, count);
    return check_launch("narrow kernel");
}

static int download_from(double* dst, const double* src, size_t count, double*, cudaStream_t st) {
    NB_CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    return NB_OK;
}
static int download_from(double* dst, const float* src, size_t count, double* staging, cudaStream_t st) {
    widen_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(src, staging, count);
    NB_TRY(check_launch("widen kernel"));
    NB_CUDA_OK(cudaMemcpyAsync(dst, staging, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    return NB_OK;
}

template <typename T> struct Abi;
template <> struct Abi<double> {
    static constexpr int is_f64 = 1;
    static int pack(const double* p, const void* m, int f, int n, double* s, cudaStream_t st) { return nb_pack_f64(p, m, f, n, s, st); }
    static int unpack(const double* s, int n, double* p, cudaStream_t st) { return nb_unpack_f64(s, n, p, st); }
    static int accel(const double* s, int n, double eps, double* a, void* ws, size_t wb, cudaStream_t st) { return nb_accel_f64(s, n, 0, n, eps, a, ws, wb, st); }
    static int run(double* sa, double* sb, double* v, double* a, int n, double dt, double eps, int ns, int si, double* sp, double* sv, double* sc, void* ws, size_t wb, int* fin, cudaStream_t st) {
        return nb_run_f64(sa, sb, v, a, n, dt, eps, ns, si, sp, sv, sc, ws, wb, fin, st);
    }
};
template <> struct Abi<float> {
    static constexpr int is_f64 = 0;
    static int pack(const double* p, const void* m, int f, int n, float* s, cudaStream_t st) { return nb_pack_f32(p, m, f, n, s, st); }
    static int unpack(const float* s, int n, double* p, cudaStream_t st) { return nb_unpack_f32(s, n, p, st); }
    static int accel(const float* s, int n, double eps, float* a, void* ws, size_t wb, cudaStream_t st) { return nb_accel_f32(s, n, 0, n, eps, a, ws, wb, st); }
    static int run(float* sa, float* sb, float* v, float* a, int n, double dt, double eps, int ns, int si, double* sp, double* sv, double* sc, void* ws, size_t wb, int* fin, cudaStream_t st) {
        return nb_run_f32(sa, sb, v, a, n, dt, eps, ns, si, sp, sv, sc, ws, wb, fin, st);
    }
};

static int upload_masses(DeviceArena& ar, const void* masses, int masses_are_f32, size_t count, void** d_m) {
    const size_t bytes = count * (masses_are_f32 ? sizeof(float) : sizeof(double));
    char* p = nullptr;
    NB_TRY(ar.alloc(&p, bytes));
    NB_CUDA_OK(cudaMemcpyAsync(p, masses, bytes, cudaMemcpyHostToDevice, ar.stream));
    *d_m = p;
    return NB_OK;
}

template <typename T>
static int host_accel(const double* pos, const void* masses, int masses_are_f32, int n, double softening, double* acc) {
    NB_REQUIRE(pos && masses && acc && n > 0, "nbh_accel_direct: bad argument");
    DeviceArena ar;
    NB_CUDA_OK(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
    const size_t n3 = (size_t)n * 3;
    const int n_pad = nb_padded_bodies(n);
    double *d_pos = nullptr, *d_stage = nullptr;
    T *d_stream = nullptr, *d_acc = nullptr;
    void *d_m = nullptr, *d_ws = nullptr;
    const size_t wb = nb_workspace_bytes(n, n, Abi<T>::is_f64);
    NB_TRY(ar.alloc(&d_pos, n3));
    NB_TRY(ar.alloc(&d_stage, n3));
    NB_TRY(ar.alloc(&d_stream, (size_t)n_pad * 4));
    NB_TRY(ar.alloc(&d_acc, n3));
    NB_TRY(ar.alloc((char**)&d_ws, wb));
    NB_CUDA_OK(cudaMemsetAsync(d_ws, 0, wb, ar.stream));  // i-tile arrival counters start at zero
    NB_TRY(upload_masses(ar, masses, masses_are_f32, n, &d_m));
    NB_CUDA_OK(cudaMemcpyAsync(d_pos, pos, n3 * sizeof(double), cudaMemcpyHostToDevice, ar.stream));
    NB_TRY(Abi<T>::pack(d_pos, d_m, masses_are_f32, n, d_stream, ar.stream));
    NB_TRY(Abi<T>::accel(d_stream, n, softening, d_acc, d_ws, wb, ar.stream));
    NB_TRY(download_from(acc, d_acc, n3, d_stage, ar.stream));
    NB_CUDA_OK(cudaStreamSynchronize(ar.stream));
    return NB_OK;
}

template <typename T>
static int host_run(double* pos, double* vel, double* acc, const void* masses, int masses_are_f32, int n, double dt,
                    double softening, int n_steps, int save_interval, double* snap_pos, double* snap_vel,
                    double* snap_acc) {
    NB_REQUIRE(pos && vel && acc && masses && n > 0, "nbh_run: bad argument");
    NB_REQUIRE(n_steps >= 0 && save_interval >= 1, "n_steps >= 0 and save_interval >= 1 required");
    DeviceArena ar;
    NB_CUDA_OK(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
    const size_t n3 = (size_t)n * 3;
    const size_t n_snap = 1 + (size_t)(n_steps / save_interval);
    const int n_pad = nb_padded_bodies(n);
    const bool snaps = snap_pos || snap_vel || snap_acc;
    double *d_pos = nullptr, *d_stage = nullptr, *d_sp = nullptr, *d_sv = nullptr, *d_sa = nullptr;
    T *d_a = nullptr, *d_b = nullptr, *d_vel = nullptr, *d_acc = nullptr;
    void *d_m = nullptr, *d_ws = nullptr;
    const size_t wb = nb_workspace_bytes(n, n, Abi<T>::is_f64);
    NB_TRY(ar.alloc(&d_pos, n3));
    NB_TRY(ar.alloc(&d_stage, n3));
    NB_TRY(ar.alloc(&d_a, (size_t)n_pad * 4));
    NB_TRY(ar.alloc(&d_b, (size_t)n_pad * 4));
    NB_TRY(ar.alloc(&d_vel, n3));
    NB_TRY(ar.alloc(&d_acc, n3));
    NB_TRY(ar.alloc((char**)&d_ws, wb));
    NB_CUDA_OK(cudaMemsetAsync(d_ws, 0, wb, ar.stream));  // i-tile arrival counters start at zero
    if (snaps) {
        NB_TRY(ar.alloc(&d_sp, n_snap * n3));
        NB_TRY(ar.alloc(&d_sv, n_snap * n3));
        NB_TRY(ar.alloc(&d_sa, n_snap * n3));
    }
    NB_TRY(upload_masses(ar, masses, masses_are_f32, n, &d_m));
    NB_CUDA_OK(cudaMemcpyAsync(d_pos, pos, n3 * sizeof(double), cudaMemcpyHostToDevice, ar.stream));
    NB_TRY(Abi<T>::pack(d_pos, d_m, masses_are_f32, n, d_a, ar.stream));
    NB_TRY(Abi<T>::pack(d_pos, d_m, masses_are_f32, n, d_b, ar.stream));
    NB_TRY(upload_as<T>(d_vel, vel, n3, d_stage, ar.stream));
    NB_TRY(upload_as<T>(d_acc, acc, n3, d_stage, ar.stream));
    int final_in_a = 1;
    NB_TRY(Abi<T>::run(d_a, d_b, d_vel, d_acc, n, dt, softening, n_steps, save_interval, d_sp, d_sv, d_sa, d_ws, wb,
                       &final_in_a, ar.stream));
    NB_TRY(Abi<T>::unpack(final_in_a ? d_a : d_b, n, d_pos, ar.stream));
    NB_CUDA_OK(cudaMemcpyAsync(pos, d_pos, n3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    NB_TRY(download_from(vel, d_vel, n3, d_stage, ar.stream));
    NB_CUDA_OK(cudaStreamSynchronize(ar.stream));  // d_stage is reused by the next narrowing download
    NB_TRY(download_from(acc, d_acc, n3, d_stage, ar.stream));
    if (snap_pos) NB_CUDA_OK(cudaMemcpyAsync(snap_pos, d_sp, n_snap * n3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    if (snap_vel) NB_CUDA_OK(cudaMemcpyAsync(snap_vel, d_sv, n_snap * n3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    if (snap_acc) NB_CUDA_OK(cudaMemcpyAsync(snap_acc, d_sa, n_snap * n3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    NB_CUDA_OK(cudaStreamSynchronize(ar.stream));
    return NB_OK;
}

}  // namespace nb

extern "C" {

int nbh_accel_direct(const double* pos, const void* masses, int masses_are_f32, int n, double softening, int use_f32,
                     double* acc) {
    return use_f32 ? nb::host_accel<float>(pos, masses, masses_are_f32, n, softening, acc)
                   : nb::host_accel<double>(pos, masses, masses_are_f32, n, softening, acc);
}

int nbh_run(double* pos, double* vel, double* acc, const void* masses, int masses_are_f32, int n, double dt,
            double softening, int n_steps, int save_interval, int use_f32, double* snap_pos, double* snap_vel,
            double* snap_acc) {
    return use_f32 ? nb::host_run<float>(pos, vel, acc, masses, masses_are_f32, n, dt, softening, n_steps,
                                         save_interval, snap_pos, snap_vel, snap_acc)
                   : nb::host_run<double>(pos, vel, acc, masses, masses_are_f32, n, dt, softening, n_steps,
                                          save_interval, snap_pos, snap_vel, snap_acc);
}

int nbh_ensemble_run(double* x, double* v, double* a, const void* masses, int masses_are_f32, int mass_stride, int B,
                     int N, double dt, double softening, int n_steps, int save_interval, int use_f32, double* out_x,
                     double* out_v, double* out_a) {
    NB_REQUIRE(x && v && a && masses && B > 0 && N > 0, "nbh_ensemble_run: bad argument");
    NB_REQUIRE(n_steps >= 0 && save_interval >= 1, "n_steps >= 0 and save_interval >= 1 required");
    nb::DeviceArena ar;
    NB_CUDA_OK(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
    const size_t bn3 = (size_t)B * N * 3;
    const int n_snap = 1 + n_steps / save_interval;
    double *d_x = nullptr, *d_v = nullptr, *d_a = nullptr, *d_ox = nullptr, *d_ov = nullptr, *d_oa = nullptr;
    void *d_m = nullptr, *d_ws = nullptr;
    const size_t wb = nb_ensemble_workspace_bytes(B);
    NB_TRY(ar.alloc(&d_x, bn3));
    NB_TRY(ar.alloc(&d_v, bn3));
    NB_TRY(ar.alloc(&d_a, bn3));
    NB_TRY(ar.alloc((char**)&d_ws, wb));
    const bool snaps = out_x || out_v || out_a;
    if (snaps) {
        NB_TRY(ar.alloc(&d_ox, bn3 * n_snap));
        NB_TRY(ar.alloc(&d_ov, bn3 * n_snap));
        NB_TRY(ar.alloc(&d_oa, bn3 * n_snap));
    }
    NB_TRY(nb::upload_masses(ar, masses, masses_are_f32, mass_stride ? (size_t)B * N : (size_t)N, &d_m));
    NB_CUDA_OK(cudaMemcpyAsync(d_x, x, bn3 * sizeof(double), cudaMemcpyHostToDevice, ar.stream));
    NB_CUDA_OK(cudaMemcpyAsync(d_v, v, bn3 * sizeof(double), cudaMemcpyHostToDevice, ar.stream));
    NB_CUDA_OK(cudaMemsetAsync(d_a, 0, bn3 * sizeof(double), ar.stream));
    auto fn = use_f32 ? nb_ensemble_f32 : nb_ensemble_f64;
    NB_TRY(fn(d_x, d_v, d_a, d_m, masses_are_f32, mass_stride, B, N, dt, softening, n_steps, save_interval,
              /*compute_a0=*/1, /*write_initial=*/1, d_ox, d_ov, d_oa, n_snap, 0, d_ws, wb, ar.stream));
    NB_CUDA_OK(cudaMemcpyAsync(x, d_x, bn3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    NB_CUDA_OK(cudaMemcpyAsync(v, d_v, bn3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    NB_CUDA_OK(cudaMemcpyAsync(a, d_a, bn3 * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    if (out_x) NB_CUDA_OK(cudaMemcpyAsync(out_x, d_ox, bn3 * n_snap * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    if (out_v) NB_CUDA_OK(cudaMemcpyAsync(out_v, d_ov, bn3 * n_snap * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    if (out_a) NB_CUDA_OK(cudaMemcpyAsync(out_a, d_oa, bn3 * n_snap * sizeof(double), cudaMemcpyDeviceToHost, ar.stream));
    NB_CUDA_OK(cudaStreamSynchronize(ar.stream));
    return NB_OK;
}

int nbh_total_energy(const double* pos, const double* vel, const void* masses, int masses_are_f32, int n,
                     double softening, double* out_kut) {
    NB_REQUIRE(pos && vel && masses && out_kut && n > 0, "nbh_total_energy: bad argument");
    nb::DeviceArena ar;
    NB_CUDA_OK(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
    const size_t n3 = (size_t)n * 3;
    double *d_pos = nullptr, *d_vel = nullptr, *d_ku = nullptr;
    void *d_m = nullptr, *d_ws = nullptr;
    const size_t wb = nb_energy_workspace_bytes(n, n);
    NB_TRY(ar.alloc(&d_pos, n3));
    NB_TRY(ar.alloc(&d_vel, n3));
    NB_TRY(ar.alloc(&d_ku, 2));
    NB_TRY(ar.alloc((char**)&d_ws, wb));
    NB_TRY(nb::upload_masses(ar, masses, masses_are_f32, n, &d_m));
    NB_CUDA_OK(cudaMemcpyAsync(d_pos, pos, n3 * sizeof(double), cudaMemcpyHostToDevice, ar.stream));
    NB_CUDA_OK(cudaMemcpyAsync(d_vel, vel, n3 * sizeof(double), cudaMemcpyHostToDevice, ar.stream));
    NB_TRY(nb_energy_f64(d_pos, d_vel, d_m, masses_are_f32, n, 0, n, softening, d_ku, d_ws, wb, ar.stream));
    double ku[2] = {0.0, 0.0};
    NB_CUDA_OK(cudaMemcpyAsync(ku, d_ku, sizeof(ku), cudaMemcpyDeviceToHost, ar.stream));
    NB_CUDA_OK(cudaStreamSynchronize(ar.stream));
    out_kut[0] = ku[0];
    out_kut[1] = ku[1];
    out_kut[2] = ku[0] + ku[1];
    return NB_OK;
}

}  // extern "C"

// Strided device -> host copy of a block of snapshot rows: `height` rows of `width` bytes, row starts
// `spitch` / `dpitch` bytes apart.  Lets the host side drain the rows of one step-chunk of every
// system (a (B, rows, N, 3) sub-block of the (B, T+1, N, 3) stacks) while the next chunk computes.
extern "C" int nb_copy_rows_d2h_async(void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t width,
                                       size_t height, nb_stream_t s) {
    NB_REQUIRE(dst_host && src_dev && width > 0 && height > 0 && dpitch >= width && spitch >= width,
               "nb_copy_rows_d2h_async: bad argument");
    NB_CUDA_OK(cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width, height, cudaMemcpyDeviceToHost,
                                 (cudaStream_t)s));
    return NB_OK;
}
