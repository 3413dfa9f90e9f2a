// rs_raster_fwd: front-to-back alpha compositing of the sorted per-tile splat lists.
// Replaces csrc/RasterizeToPixels3DGSFwd.cu:17-187 (host side csrc/Rasterization.cpp:20-115).
//
// One CTA (256 threads) per 16x16 tile, like the reference, but:
//   * each warp owns an 8x4 pixel sub-block and, per batch of 256 staged splats, tests 32 splats at a time (one per
//     lane) against its sub-block with a conservative ellipse bounding box derived from the conic and the opacity
//     (alpha >= 1/255  <=>  sigma <= ln(255 * opacity)); a ballot gives the list of splats that can touch the
//     sub-block and only those are evaluated.  A skipped (pixel, splat) pair is one the reference would have
//     `continue`d on (alpha < 1/255), so results are unchanged while ~2/3 of the evaluations disappear.
//   * colours are staged in shared memory with the geometry (the reference re-reads them from global per pixel).
//   * a warp whose 32 pixels are saturated stops evaluating (the reference only stops per CTA).
//   * any channel count is handled (template capacity >= channels, > 32 in chunks) -- no python-side padding.
// The per-pixel arithmetic reproduces the reference's compiled instruction sequence (FMUL/FFMA association read from
// its SASS: sigma = fma(dy, b*dx, 0.5 * fma(dx, a*dx, (c*dy)*dy)), alpha = min(.999, op * ex2(-sigma*log2e)), FTZ) so
// that images, alphas and last_ids are bit-identical to the reference on identical inputs.
// This stage is issue-bound (FP32 + MUFU), not HBM-bound.
#include "common.cuh"

#define RAST_THREADS 256

template <int CDIM> struct RastSmem {
    float4 xyoa[RAST_THREADS]; // x, y, opacity, conic.a
    float4 bcee[RAST_THREADS]; // conic.b, conic.c, half-extent x, half-extent y (cull box)
    float color[CDIM][RAST_THREADS]; // channel-major: conflict-free staging, broadcast reads
};

// conservative half extents of {p : sigma(p) <= ln(255 op)}; 3e38 = "cannot cull", -3e38 = "can never contribute"
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

template <int CDIM>
__global__ void __launch_bounds__(RAST_THREADS)
rs_raster_fwd_kernel(const rs_raster_fwd_args a, const int ch_off, const int ch_cnt) {
    __shared__ RastSmem<CDIM> sm;

    const uint32_t tiles_per_image = (uint32_t)(a.tile_width * a.tile_height);
    const uint32_t image_id = blockIdx.x / tiles_per_image;
    const uint32_t tile_id = blockIdx.x - image_id * tiles_per_image;
    const uint32_t tile_y = tile_id / (uint32_t)a.tile_width;
    const uint32_t tile_x = tile_id - tile_y * (uint32_t)a.tile_width;

    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    // warp -> 8x4 pixel sub-block, lane -> pixel inside it
    const uint32_t sub_x = tile_x * RS_TILE + (warp & 1) * 8;
    const uint32_t sub_y = tile_y * RS_TILE + (warp >> 1) * 4;
    const uint32_t j = sub_x + (lane & 7);
    const uint32_t i = sub_y + (lane >> 3);
    const float px = (float)j + 0.5f;
    const float py = (float)i + 0.5f;
    const bool inside = (i < (uint32_t)a.image_height && j < (uint32_t)a.image_width);
    const size_t pix_id = (size_t)image_id * a.image_height * a.image_width + (size_t)i * a.image_width + j;

    const float *bg = a.backgrounds != nullptr ? a.backgrounds + (size_t)image_id * a.channels + ch_off : nullptr;

    // RasterizeToPixels3DGSFwd.cu:73-80: masked-out tile -> background colour only, alphas / last_ids untouched
    if (a.masks != nullptr && !a.masks[(size_t)image_id * tiles_per_image + tile_id]) {
        if (inside) {
            for (int k = 0; k < ch_cnt; ++k)
                a.render_colors[pix_id * a.channels + ch_off + k] = bg == nullptr ? 0.0f : bg[k];
        }
        return;
    }

    const int64_t n_isects = a.n_isects_dev != nullptr ? min((int64_t)*a.n_isects_dev, a.n_isects) : a.n_isects;
    const int32_t *offs = a.tile_offsets + (size_t)image_id * tiles_per_image;
    const int32_t range_start = offs[tile_id];
    const int32_t range_end = (image_id == (uint32_t)a.I - 1 && tile_id == tiles_per_image - 1)
                                  ? (int32_t)n_isects
                                  : offs[tile_id + 1];
    const int num_batches = (range_end - range_start + RAST_THREADS - 1) / RAST_THREADS;

    // sub-block bounds in pixel-centre coordinates
    const float bx0 = (float)sub_x + 0.5f, bx1 = (float)sub_x + 7.5f;
    const float by0 = (float)sub_y + 0.5f, by1 = (float)sub_y + 3.5f;

    float T = 1.0f;
    uint32_t cur_idx = 0;
    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);
    float pix_out[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k)
        pix_out[k] = 0.f;

    for (int b = 0; b < num_batches; ++b) {
        // also the barrier that protects the staging buffers from the previous batch
        if (__syncthreads_count(done) >= RAST_THREADS)
            break;

        const int32_t batch_start = range_start + RAST_THREADS * b;
        const int32_t idx = batch_start + tr;
        if (idx < range_end) {
            const int32_t g = a.flatten_ids[idx];
            const float2 xy = reinterpret_cast<const float2 *>(a.means2d)[g];
            const int32_t go = a.attr_mod_opacities > 0 ? g % a.attr_mod_opacities : g;
            const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
            const float op = a.opacities[go];
            const float ca = a.conics[(size_t)g * 3 + 0];
            const float cb = a.conics[(size_t)g * 3 + 1];
            const float cc = a.conics[(size_t)g * 3 + 2];
            float ex, ey;
            rs_cull_extents(ca, cb, cc, op, ex, ey);
            sm.xyoa[tr] = make_float4(xy.x, xy.y, op, ca);
            sm.bcee[tr] = make_float4(cb, cc, ex, ey);
            const float *cp = a.colors + (size_t)gc * a.channels + ch_off;
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                if (k < ch_cnt)
                    sm.color[k][tr] = cp[k];
        }
        __syncthreads();

        if (!warp_done) {
            const int batch_size = min(RAST_THREADS, range_end - batch_start);
            for (int chunk = 0; chunk < batch_size; chunk += 32) {
                const int t = chunk + lane;
                bool hit = false;
                if (t < batch_size) {
                    const float4 g0 = sm.xyoa[t];
                    const float4 g1 = sm.bcee[t];
                    hit = (g0.x + g1.z >= bx0) && (g0.x - g1.z <= bx1) && (g0.y + g1.w >= by0) &&
                          (g0.y - g1.w <= by1);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int tt = chunk + __ffs(m) - 1;
                    m &= m - 1;
                    if (!done) {
                        const float4 g0 = sm.xyoa[tt];
                        const float4 g1 = sm.bcee[tt];
                        const float dx = __fsub_rn(g0.x, px);
                        const float dy = __fsub_rn(g0.y, py);
                        const float tc = __fmul_rn(__fmul_rn(g1.y, dy), dy);
                        const float s = __fmaf_rn(dx, __fmul_rn(g0.w, dx), tc);
                        const float sigma = __fmaf_rn(dy, __fmul_rn(g1.x, dx), __fmul_rn(s, 0.5f));
                        const float alpha = fminf(0.999f, __fmul_rn(g0.z, __expf(-sigma)));
                        if (!(sigma < 0.f || alpha < RS_ALPHA_THRESHOLD)) {
                            const float next_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                            if (next_T <= 1e-4f) {
                                done = true;
                            } else {
                                const float vis = __fmul_rn(alpha, T);
#pragma unroll
                                for (int k = 0; k < CDIM; ++k)
                                    if (k < ch_cnt)
                                        pix_out[k] = __fmaf_rn(sm.color[k][tt], vis, pix_out[k]);
                                cur_idx = (uint32_t)(batch_start + tt);
                                T = next_T;
                            }
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    break;
                }
            }
        }
    }

    if (inside) {
        a.render_alphas[pix_id] = __fsub_rn(1.0f, T);
        float *out = a.render_colors + pix_id * a.channels + ch_off;
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < ch_cnt)
                out[k] = bg == nullptr ? pix_out[k] : __fmaf_rn(T, bg[k], pix_out[k]);
        a.last_ids[pix_id] = (int32_t)cur_idx;
    }
}

template <int CDIM>
static int launch_raster_fwd(const rs_raster_fwd_args &a, int ch_off, int ch_cnt, cudaStream_t s) {
    const int64_t grid = (int64_t)a.I * a.tile_width * a.tile_height;
    rs_raster_fwd_kernel<CDIM><<<(unsigned)grid, RAST_THREADS, 0, s>>>(a, ch_off, ch_cnt);
    RS_LAUNCH_CHECK("rs_raster_fwd_kernel");
    return 0;
}

int rs_raster_fwd_chunk(const rs_raster_fwd_args &a, int ch_off, int ch_cnt, cudaStream_t s) {
    if (ch_cnt <= 1)
        return launch_raster_fwd<1>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 2)
        return launch_raster_fwd<2>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 3)
        return launch_raster_fwd<3>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 4)
        return launch_raster_fwd<4>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 5)
        return launch_raster_fwd<5>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 8)
        return launch_raster_fwd<8>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 16)
        return launch_raster_fwd<16>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 17)
        return launch_raster_fwd<17>(a, ch_off, ch_cnt, s);
    return launch_raster_fwd<32>(a, ch_off, ch_cnt, s);
}

int rs_check_raster_args(const rs_raster_fwd_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->tile_size == RS_TILE, "%s: tile_size must be 16 (got %d)", who, a->tile_size);
    RS_CHECK(a->channels >= 1 && a->channels <= RS_MAX_CHANNELS, "%s: Unsupported number of color channels: %d", who,
             a->channels);
    RS_CHECK(a->I >= 0 && a->image_width > 0 && a->image_height > 0, "%s: bad image geometry", who);
    RS_CHECK(a->tile_width * a->tile_size >= a->image_width && a->tile_height * a->tile_size >= a->image_height,
             "%s: tile grid %dx%d does not cover the %dx%d image", who, a->tile_width, a->tile_height, a->image_width,
             a->image_height);
    RS_CHECK((int64_t)a->I * a->tile_width * a->tile_height < ((int64_t)1 << 31), "%s: too many tiles", who);
    RS_CHECK(a->n_isects >= 0 && a->n_isects < ((int64_t)1 << 31), "%s: n_isects out of range", who);
    return 0;
}

extern "C" int rs_raster_fwd(const rs_raster_fwd_args *a, rs_stream_t stream) {
    if (int e = rs_check_raster_args(a, "rs_raster_fwd"))
        return e;
    if (a->I == 0)
        return 0;
    RS_CHECK(a->means2d && a->conics && a->colors && a->opacities && a->tile_offsets && a->render_colors &&
                 a->render_alphas && a->last_ids,
             "rs_raster_fwd: null pointer");
    RS_CHECK(a->n_isects == 0 || a->flatten_ids != nullptr, "rs_raster_fwd: null flatten_ids");
    for (int off = 0; off < a->channels; off += 32) {
        const int cnt = a->channels - off < 32 ? a->channels - off : 32;
        if (int e = rs_raster_fwd_chunk(*a, off, cnt, (cudaStream_t)stream))
            return e;
    }
    return 0;
}
