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

// Tight tile list of one splat (rs_project_fwd_args.tile_footprints): the tiles of the reference's bounding rectangle that
// hold a pixel centre where the splat can reach alpha >= 1/255 -- the same test, with the same limit, the compositing
// kernel applies per warp (a warp's pixel rows lie inside the tile, so a tile that fails here fails in every warp).
// Returns the number of listed tiles; rectangles of more than 64 tiles keep all their tiles (mask = all ones).
__device__ __forceinline__ int rs_tile_footprint(float mx, float my, int32_t rx, int32_t ry, float qa, float qb, float qc,
                                                 float opac, uint32_t tile_size, uint32_t tile_width,
                                                 uint32_t tile_height, uint4 &fp) {
    fp = make_uint4(0u, 0u, 0u, 0u);
    if (rx <= 0 || ry <= 0)
        return 0;
    const RsTileRect r = rs_tile_rect(mx, my, (float)rx, (float)ry, tile_size, tile_width, tile_height);
    const uint32_t w = r.x1 - r.x0, h = r.y1 - r.y0, n = w * h;
    unsigned long long mask = ~0ull;
    int cnt = (int)n;
    if (n > 0u && n <= 64u) {
        const float limit = rs_cull_limit(qa, qb, qc, opac);
        const float ts = (float)tile_size;
        mask = 0ull;
        uint32_t t = 0;
        for (uint32_t ty = r.y0; ty < r.y1; ++ty) {
            const float y0 = (float)ty * ts + 0.5f, y1 = y0 + ts - 1.f; // pixel centres of the tile
            for (uint32_t tx = r.x0; tx < r.x1; ++tx, ++t) {
                const float x0 = (float)tx * ts + 0.5f, x1 = x0 + ts - 1.f;
                if (rs_splat_touches_rect(mx, my, qa, qb, qc, limit, x0, x1, y0, y1))
                    mask |= 1ull << t;
            }
        }
        cnt = __popcll(mask);
    }
    fp = make_uint4((uint32_t)mask, (uint32_t)(mask >> 32), r.x0 | (r.y0 << 16), w | (h << 16));
    return cnt;
}

// The same footprint computed by a whole warp for its 32 splats: the tiles of the 32 bounding rectangles are numbered
// consecutively and dealt out to the lanes, so a warp spends sum(tiles) / 32 rounds instead of max(tiles) -- the per-lane
// loop above costs the projection kernel +70 % on the 1 M-Gaussian scene, this one a few per cent.  Must be called by all
// 32 lanes (valid = false for a lane without a splat); `fs` is the calling warp's scratch.
struct RsFootWarp {
    float4 p0[32]; // centre x, y, conic a, b
    float4 p1[32]; // conic c, cull limit, bits(x0 | y0 << 16), bits(w)
    int excl[32];
    unsigned int mask[32][2];
};
__device__ __forceinline__ int rs_tile_footprint_warp(bool valid, float mx, float my, int32_t rx, int32_t ry, float qa,
                                                      float qb, float qc, float opac, uint32_t tile_size,
                                                      uint32_t tile_width, uint32_t tile_height, uint4 &fp,
                                                      RsFootWarp &fs) {
    const int lane = threadIdx.x & 31;
    uint32_t x0 = 0, y0 = 0, w = 0, h = 0, n = 0;
    if (valid && rx > 0 && ry > 0) {
        const RsTileRect r = rs_tile_rect(mx, my, (float)rx, (float)ry, tile_size, tile_width, tile_height);
        x0 = r.x0;
        y0 = r.y0;
        w = r.x1 - r.x0;
        h = r.y1 - r.y0;
        n = w * h;
    }
    const bool masked = n > 0u && n <= 64u;
    const int m = masked ? (int)n : 0;
    int incl = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    // the lanes that own tiles, compacted: owner c = the c-th such lane
    const unsigned has = __ballot_sync(0xffffffffu, m > 0);
    const int ci = __popc(has & ((1u << lane) - 1u));
    if (m > 0) {
        fs.excl[ci] = incl - m;
        fs.p0[ci] = make_float4(mx, my, qa, qb);
        // (65536 / w + 1) * t >> 16 = t / w for t < 64, w <= 64
        fs.p1[ci] = make_float4(qc, rs_cull_limit(qa, qb, qc, opac), __uint_as_float(x0 | (y0 << 16)),
                                __uint_as_float((65536u / w + 1u) | (w << 20)));
        fs.mask[ci][0] = 0u;
        fs.mask[ci][1] = 0u;
    }
    __syncwarp();
    const float ts = (float)tile_size;
    int started = 0; // owners whose first tile lies before the current window of 32 tiles
    for (int base = 0; base < total; base += 32) {
        const int rel = incl - m - base;
        const unsigned heads = __reduce_or_sync(0xffffffffu, (m > 0 && rel >= 0 && rel < 32) ? 1u << rel : 0u);
        const int j = base + lane;
        if (j < total) {
            const int oc = started + __popc(heads & (0xffffffffu >> (31 - lane))) - 1; // owner of tile j
            const uint32_t t = (uint32_t)(j - fs.excl[oc]);
            const float4 g0 = fs.p0[oc], g1 = fs.p1[oc];
            const uint32_t xy = __float_as_uint(g1.z), iw = __float_as_uint(g1.w);
            const uint32_t row = (t * (iw & 0xfffffu)) >> 16;
            const uint32_t col = t - row * (iw >> 20);
            const float fx0 = (float)((xy & 0xffffu) + col) * ts + 0.5f, fy0 = (float)((xy >> 16) + row) * ts + 0.5f;
            if (rs_splat_touches_rect(g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, fx0, fx0 + ts - 1.f, fy0, fy0 + ts - 1.f))
                atomicOr(&fs.mask[oc][t >> 5], 1u << (t & 31u));
        }
        started += __popc(heads);
    }
    __syncwarp();
    unsigned long long mask = ~0ull;
    int cnt = (int)n;
    if (masked) {
        mask = (unsigned long long)fs.mask[ci][0] | ((unsigned long long)fs.mask[ci][1] << 32);
        cnt = __popcll(mask);
    }
    __syncwarp(); // the scratch may be reused by the caller's next round
    fp = n > 0u ? make_uint4((uint32_t)mask, (uint32_t)(mask >> 32), x0 | (y0 << 16), w | (h << 16)) : make_uint4(0u, 0u, 0u, 0u);
    return cnt;
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
