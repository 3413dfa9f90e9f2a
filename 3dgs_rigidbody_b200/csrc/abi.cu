// C-ABI plumbing: error string, version, struct sizes, device properties.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void rs_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *rs_last_error(void) { return g_err; }

static uint64_t g_launches = 0;

// Per-kernel timing of whatever this thread launches between rs_profile_begin() and rs_profile_end(): one CUDA event is
// recorded on the profiled stream behind every kernel launch (RS_LAUNCH_CHECK), so the time between consecutive events is
// the duration of one kernel as it ran in stream order (including its launch gap; a short delay kernel in front lets the
// host run ahead so the gaps are the GPU's, not the host's).  Measurement aid for bench.py.
#define RS_PROFILE_MAX 256
struct RsProfile {
    bool on;
    cudaStream_t stream;
    int n;                               // kernels recorded
    cudaEvent_t ev[RS_PROFILE_MAX + 1];  // ev[0] = begin
    const char *name[RS_PROFILE_MAX];
};
static thread_local RsProfile g_prof = {false, nullptr, 0, {nullptr}, {nullptr}};

void rs_count_launch(const char *name) {
    __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED);
    if (g_prof.on && g_prof.n < RS_PROFILE_MAX) {
        const int i = g_prof.n + 1;
        if (g_prof.ev[i] == nullptr && cudaEventCreate(&g_prof.ev[i]) != cudaSuccess)
            return;
        if (cudaEventRecord(g_prof.ev[i], g_prof.stream) == cudaSuccess) {
            g_prof.name[g_prof.n] = name;
            g_prof.n = i;
        }
    }
}

// Holds the stream for ~0.5 ms so that the host can enqueue the whole profiled frame (kernels + events) before the GPU
// starts on it: the event-to-event times are then kernel times, not the host's launch cadence.
__global__ void rs_profile_delay_kernel(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {
    }
}

extern "C" int rs_profile_begin(rs_stream_t stream) {
    g_prof.stream = (cudaStream_t)stream;
    g_prof.n = 0;
    if (g_prof.ev[0] == nullptr)
        RS_CUDA(cudaEventCreate(&g_prof.ev[0]));
    rs_profile_delay_kernel<<<1, 1, 0, g_prof.stream>>>(1000000ll);
    RS_CUDA(cudaGetLastError());
    RS_CUDA(cudaEventRecord(g_prof.ev[0], g_prof.stream));
    g_prof.on = true;
    return 0;
}

extern "C" int rs_profile_end(int32_t max_kernels, float *ms, const char **names, int32_t *n_kernels) {
    RS_CHECK(g_prof.on, "rs_profile_end: rs_profile_begin was not called on this thread");
    g_prof.on = false;
    RS_CHECK(ms != nullptr && names != nullptr && n_kernels != nullptr && max_kernels >= 0, "rs_profile_end: bad arguments");
    const int n = g_prof.n < max_kernels ? g_prof.n : max_kernels;
    if (g_prof.n > 0)
        RS_CUDA(cudaEventSynchronize(g_prof.ev[g_prof.n]));
    for (int i = 0; i < n; ++i) {
        RS_CUDA(cudaEventElapsedTime(&ms[i], g_prof.ev[i], g_prof.ev[i + 1]));
        names[i] = g_prof.name[i];
    }
    *n_kernels = n;
    return 0;
}
extern "C" uint64_t rs_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int rs_abi_version(void) { return RS_ABI_VERSION; }

extern "C" uint64_t rs_sizeof_args(int which) {
    switch (which) {
    case 0:
        return sizeof(rs_project_fwd_args);
    case 1:
        return sizeof(rs_project_bwd_args);
    case 2:
        return sizeof(rs_isect_args);
    case 3:
        return sizeof(rs_sort_args);
    case 4:
        return sizeof(rs_raster_fwd_args);
    case 5:
        return sizeof(rs_raster_bwd_args);
    case 6:
        return sizeof(rs_frame_args);
    case 7:
        return sizeof(rs_rigid_t);
    case 8:
        return sizeof(rs_isect_sorted_args);
    case 9:
        return sizeof(rs_sh_args);
    case 10:
        return sizeof(rs_project_packed_fwd_args);
    case 11:
        return sizeof(rs_exchange_args);
    case 12:
        return sizeof(rs_cgc_args);
    case 13:
        return sizeof(rs_seghead_args);
    case 14:
        return sizeof(rs_exchange_grad_args);
    default:
        return 0;
    }
}

int rs_current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return -1;
    return dev;
}

int rs_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
