// nb_core.cu -- error state, device info and the host-side planning arithmetic of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "nb_common.cuh"

namespace nb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return NB_ERR_CUDA;
}

}  // namespace nb

extern "C" {

int nb_abi_version(void) { return NB_ABI_VERSION; }

const char* nb_last_error(void) { return nb::g_error; }

int nb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz) {
    int v = 0;
    if (sm_count) {
        NB_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        *sm_count = v;
    }
    if (cc_major) {
        NB_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
        *cc_major = v;
    }
    if (cc_minor) {
        NB_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
        *cc_minor = v;
    }
    if (sm_clock_khz) {
        NB_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device));
        *sm_clock_khz = v;
    }
    return NB_OK;
}

int nb_padded_bodies(int n) {
    if (n <= 0) return nb::kChunkBodies;
    return nb::round_up(n, nb::kChunkBodies);
}

// The j axis is cut into n_seg segments of seg_len bodies (the last may be shorter).  The plan is
// a function of n alone, so a body's acceleration is the same bits whether it is evaluated by a
// one-GPU launch or inside any rank's i-slab.  Policy: segments of n/16 bodies clamped to
// [256, 4096] (measured flat from 1024 up at N = 65,536; longer segments mean fewer partials for
// the finish pass), at most 64 segments -- enough CTAs (i-tiles x segments) to fill 148 SMs several
// times over from n ~ 16k up, even when eight ranks each take an eighth of the i-tiles.
int nb_segment_plan(int n, int* seg_len, int* n_seg) {
    const int n_pad = nb_padded_bodies(n);
    int len = nb::round_up(nb::ceil_div(n_pad, 16), nb::kChunkBodies);
    if (len < 256) len = 256;  // mid-size systems (1k..16k bodies): short serial chains, more CTAs
    if (len > 4096) len = 4096;
    if (nb::ceil_div(n_pad, len) > 64) len = nb::round_up(nb::ceil_div(n_pad, 64), nb::kChunkBodies);
    if (len > n_pad) len = n_pad;
    if (seg_len) *seg_len = len;
    if (n_seg) *n_seg = nb::ceil_div(n_pad, len);
    return NB_OK;
}

size_t nb_workspace_bytes(int n, int n_i, int is_f64) {
    int n_seg = 1;
    nb_segment_plan(n, nullptr, &n_seg);
    const size_t elt = is_f64 ? sizeof(double) : sizeof(float);
    size_t bytes = (size_t)n_seg * 3 * (size_t)(n_i > 0 ? n_i : 1) * elt;
    bytes = (bytes + 255) / 256 * 256;
    // header: one arrival counter per i-tile of the WHOLE system (smallest tile: 128 bodies), so the layout does not
    // depend on the slab; must be zero before the first launch that uses the workspace -- every launch leaves it zero
    bytes += nb::ws_header_bytes(n);
    if (n_i == n) bytes += nb::ws_persist_bytes(n);  // whole-system runs may take the one-launch kernel (nb_persist.cu)
    return bytes;
}

}  // extern "C"
