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
void rs_count_launch() { __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED); }
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
    default:
        return 0;
    }
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
