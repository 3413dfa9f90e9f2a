// Shared device/host helpers for librigidsplat (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rigidsplat.h"

#define RS_ALPHA_THRESHOLD (1.f / 255.f) // gsplat/cuda/include/Common.h:54
#define RS_TILE 16                       // the only tile size the reference exercises (rendering.py:184-185)

// elements (image x gaussian pairs) covered by one CTA of the projection / tile-count / key-emission kernels;
// also the granularity of `block_sums`.
#define RS_ISECT_BLOCK 1024
#define RS_ISECT_THREADS 256

void rs_set_error(const char *fmt, ...);

#define RS_CHECK(cond, ...)                                                                                            \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            rs_set_error(__VA_ARGS__);                                                                                 \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

#define RS_CUDA(call)                                                                                                  \
    do {                                                                                                               \
        cudaError_t err__ = (call);                                                                                    \
        if (err__ != cudaSuccess) {                                                                                    \
            rs_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(err__));             \
            return 2;                                                                                                  \
        }                                                                                                              \
    } while (0)

void rs_count_launch(const char *name);
#define RS_LAUNCH_CHECK(name)                                                                                          \
    do {                                                                                                               \
        rs_count_launch(name);                                                                                         \
        cudaError_t err__ = cudaGetLastError();                                                                        \
        if (err__ != cudaSuccess) {                                                                                    \
            rs_set_error("launch of %s failed: %s", name, cudaGetErrorString(err__));                                  \
            return 3;                                                                                                  \
        }                                                                                                              \
    } while (0)

static inline int rs_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// bit width helpers matching `(uint32_t)floor(log2(x)) + 1` of csrc/Intersect.cpp:50-51
static inline __host__ __device__ uint32_t rs_bit_width(uint32_t x) {
    uint32_t n = 0;
    while (x) {
        ++n;
        x >>= 1;
    }
    return n;
}

// number of SMs of the current device (cached per device)
int rs_num_sms();

// Per-DEVICE once flags (kernel attributes such as the opt-in shared-memory size are per device, not per process):
//   static RsPerDevice f;  if (!rs_dev_done(f)) { RS_CUDA(cudaFuncSetAttribute(...)); rs_dev_mark(f); }
struct RsPerDevice {
    unsigned long long done[2]; // bit d = set on device d (128 devices; beyond that the attribute is set on every call)
};
int rs_current_device();
static inline bool rs_dev_done(const RsPerDevice &f) {
    const int d = rs_current_device();
    return d >= 0 && d < 128 && ((f.done[d >> 6] >> (d & 63)) & 1ull);
}
static inline void rs_dev_mark(RsPerDevice &f) {
    const int d = rs_current_device();
    if (d >= 0 && d < 128)
        __atomic_fetch_or(&f.done[d >> 6], 1ull << (d & 63), __ATOMIC_RELAXED);
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned rs_lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Tile rectangle of one projected Gaussian: csrc/IntersectTile.cu:55-84.
// tile_min inclusive, tile_max exclusive; float->uint conversion saturates negatives to 0 exactly like the reference's
// `(uint32_t)floor(...)`.
struct RsTileRect {
    uint32_t x0, y0, x1, y1;
};
__device__ __forceinline__ RsTileRect rs_tile_rect(float mx, float my, float rx, float ry, uint32_t tile_size,
                                                   uint32_t tile_width, uint32_t tile_height) {
    float ts = static_cast<float>(tile_size);
    float tile_radius_x = rx / ts;
    float tile_radius_y = ry / ts;
    float tile_x = mx / ts;
    float tile_y = my / ts;
    RsTileRect r;
    r.x0 = min((uint32_t)floorf(tile_x - tile_radius_x), tile_width);
    r.y0 = min((uint32_t)floorf(tile_y - tile_radius_y), tile_height);
    r.x1 = min((uint32_t)ceilf(tile_x + tile_radius_x), tile_width);
    r.y1 = min((uint32_t)ceilf(tile_y + tile_radius_y), tile_height);
    return r;
}
__device__ __forceinline__ int32_t rs_tile_count(int32_t radius_x, int32_t radius_y, float mx, float my,
                                                 uint32_t tile_size, uint32_t tile_width, uint32_t tile_height) {
    if (radius_x <= 0 || radius_y <= 0)
        return 0;
    RsTileRect r = rs_tile_rect(mx, my, (float)radius_x, (float)radius_y, tile_size, tile_width, tile_height);
    return (int32_t)((r.y1 - r.y0) * (r.x1 - r.x0));
}

// Conservative half extents (in pixels) of {p : sigma(p) <= ln(255 * opacity)}, i.e. of the region where a splat can
// reach alpha >= 1/255 (RasterizeToPixels3DGSFwd.cu:148-149 skips everything else).  3e38 = "cannot cull",
// -3e38 = "can never contribute".  Used to build the per-splat compositing record.
__device__ __forceinline__ void rs_cull_extents(float a, float b, float c, float op, float &ex, float &ey) {
    ex = 3e38f;
    ey = 3e38f;
    const float det = a * c - b * b;
    const float L = logf(op * 255.f);
    if (op < RS_ALPHA_THRESHOLD * 0.999f) { // alpha <= op < 1/255 whenever sigma >= 0
        ex = -3e38f;
        ey = -3e38f;
        return;
    }
    if (a > 0.f && c > 0.f && det > 0.f && a * c <= 256.f * det && L == L) {
        const float Lm = L + 1e-3f * (1.f + fabsf(L));
        if (Lm <= 0.f) {
            ex = 0.25f;
            ey = 0.25f;
            return;
        }
        const float inv = 2.f * Lm / det;
        const float hx = sqrtf(inv * c) * 1.0005f + 0.25f;
        const float hy = sqrtf(inv * a) * 1.0005f + 0.25f;
        if (hx < 4096.f && hy < 4096.f) {
            ex = hx;
            ey = hy;
        }
    }
}

// Cull limit of a splat for the exact ellipse-vs-rectangle test of raster_fwd.cu: a pixel can only reach alpha >= 1/255
// if sigma(pixel) <= ln(255 * opacity) (RasterizeToPixels3DGSFwd.cu:148-149); the limit carries a 1e-3 margin for the
// float error of both evaluations.  +3e38 = "cannot cull" (degenerate conic), -3e38 = "can never contribute".
__device__ __forceinline__ float rs_cull_limit(float a, float b, float c, float op) {
    if (op < RS_ALPHA_THRESHOLD * 0.999f) // alpha <= op < 1/255 whenever sigma >= 0
        return -3e38f;
    const float det = a * c - b * b;
    const float L = logf(op * 255.f);
    if (a > 0.f && c > 0.f && det > 0.f && a * c <= 256.f * det && L == L)
        return L + 1e-3f * (1.f + fabsf(L));
    return 3e38f;
}

// Exact test "can the splat reach alpha >= 1/255 somewhere in the rectangle [x0,x1] x [y0,y1] (pixel-centre coordinates)":
// minimum of sigma over the rectangle against the cull limit.  sigma is a convex quadratic centred on the splat, so the
// minimum is 0 if the centre is inside, else it lies on the edge(s) facing the centre: at most one vertical and one
// horizontal edge, each a clamped 1-D minimisation.  NaN -> true (evaluate).
__device__ __forceinline__ bool rs_splat_touches_rect(float cx, float cy, float qa, float qb, float qc, float limit,
                                                      float x0, float x1, float y0, float y1) {
    const float dx = cx - fminf(fmaxf(cx, x0), x1); // 0 when the centre is within the x range
    const float dy = cy - fminf(fmaxf(cy, y0), y1);
    // vertical edge (fixed dx): optimum dy* = -b dx / c, clamped to the edge
    const float pyv = fminf(fmaxf(cy + __fdividef(qb * dx, qc), y0), y1);
    const float d2 = cy - pyv;
    const float qv = 0.5f * (qa * dx * dx + qc * d2 * d2) + qb * dx * d2;
    // horizontal edge (fixed dy): optimum dx* = -b dy / a
    const float pxh = fminf(fmaxf(cx + __fdividef(qb * dy, qa), x0), x1);
    const float d1 = cx - pxh;
    const float qh = 0.5f * (qa * d1 * d1 + qc * dy * dy) + qb * d1 * dy;
    float qmin = 0.f;
    if (dx != 0.f)
        qmin = qv;
    if (dy != 0.f)
        qmin = (dx != 0.f) ? fminf(qv, qh) : qh;
    return !(qmin > limit);
}

// block-wide sum of one int per thread (blockDim.x == RS_ISECT_THREADS), result valid in thread 0
__device__ __forceinline__ int rs_block_sum_256(int v, int *smem8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
        smem8[threadIdx.x >> 5] = v;
    __syncthreads();
    int s = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < RS_ISECT_THREADS / 32; ++w)
            s += smem8[w];
    }
    return s;
}
#endif
