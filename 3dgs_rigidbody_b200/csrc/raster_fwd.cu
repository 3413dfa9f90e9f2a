// rs_raster_fwd: front-to-back alpha compositing of the sorted per-tile splat lists.
// Replaces csrc/RasterizeToPixels3DGSFwd.cu:17-187 (host side csrc/Rasterization.cpp:20-115).
//
// One CTA (256 threads) per 16x16 tile, like the reference, but:
//   * splats are staged as 32-byte RECORDS {x, y, opacity, conic a | conic b, conic c, cull limit, 0} that the
//     projection kernel (frame path) or a small pack kernel (operator path) writes once per (camera, Gaussian); the
//     compositing kernel copies them global -> shared with cp.async (LDGSTS, no register staging) through a ring of
//     RAST_STAGES batches of 256 splats, so the gathers of batch b+2 are in flight while batch b is composited and
//     there is ONE block barrier per batch (the reference: load, barrier, composite, barrier, nothing in flight);
//   * each warp owns an 8x4 pixel sub-block and tests 32 staged splats at a time (one per lane) against it with the
//     record's region {alpha >= 1/255} (exact ellipse-vs-rectangle test); a ballot gives the splats that can touch the sub-block
//     and only those are evaluated.  A skipped (pixel, splat) pair is one the reference would have `continue`d on
//     (alpha < 1/255), so results are unchanged while most of the evaluations disappear;
//   * colours are staged in shared memory with the geometry (the reference re-reads them from global per pixel);
//   * a warp whose 32 pixels are saturated stops evaluating (the reference only stops per CTA);
//   * any channel count is handled (template capacity >= channels, > 32 in chunks) -- no python-side padding.
// The per-pixel arithmetic reproduces the reference's compiled instruction sequence (FMUL/FFMA association read from
// its SASS: sigma = fma(dy, b*dx, 0.5 * fma(dx, a*dx, (c*dy)*dy)), alpha = min(.999, op * ex2(-sigma*log2e)), FTZ) so
// that images, alphas and last_ids are bit-identical to the reference on identical inputs.
// This stage is issue-bound (FP32 + MUFU), not HBM-bound.
#include <stdlib.h>

#include "raster_common.cuh"

#ifndef RS_RASTER_PERSISTENT_DEFAULT
#define RS_RASTER_PERSISTENT_DEFAULT 0
#endif


// records [n, 8] from the operator-level tensors (means2d, conics, opacities)
__global__ void __launch_bounds__(256) rs_raster_pack_kernel(const rs_raster_fwd_args a, const int64_t n_rows) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows)
        return;
    const float2 xy = reinterpret_cast<const float2 *>(a.means2d)[g];
    const float ca = a.conics[g * 3 + 0], cb = a.conics[g * 3 + 1], cc = a.conics[g * 3 + 2];
    const float op = a.opacities[a.attr_mod_opacities > 0 ? g % a.attr_mod_opacities : g];
    float4 *rec = reinterpret_cast<float4 *>(a.records) + g * 2;
    rec[0] = make_float4(xy.x, xy.y, op, ca);
    rec[1] = make_float4(cb, cc, rs_cull_limit(ca, cb, cc, op), 0.f);
}


// CP = colour row pitch in shared memory (floats): CDIM rounded up to a multiple of 4 so rows can be read as float4
#define RAST_BATCH 256                   // splats per ring stage
#define RAST_CONSUMERS 8                 // compositing warps (one 8x4 pixel sub-block each)
#define RAST_THREADS (32 * (RAST_CONSUMERS + 1)) // + one producer warp
#ifndef RS_RASTER_3STAGE_MAX_CDIM
#define RS_RASTER_3STAGE_MAX_CDIM 8 // widest colour row that still gets a 3-stage ring (shared memory per CTA)
#endif
template <int CDIM> struct RastCfg {
    static constexpr int CP = (CDIM + 3) & ~3;
    static constexpr int STAGES = (CDIM <= RS_RASTER_3STAGE_MAX_CDIM) ? 3 : 2;
    static constexpr int STAGE_FLOATS = RAST_BATCH * (8 + CP);
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_FLOATS * sizeof(float);
};

// Warp-specialised compositing: warp 8 is the PRODUCER -- it gathers the tile's splats (record + colour row per flatten
// id) into a ring of STAGES batches with cp.async and signals full[stage] through cp.async.mbarrier.arrive; warps 0..7 are
// CONSUMERS -- each waits on full[stage], composites its 8x4 pixels and arrives on empty[stage].  Consumers never wait
// for each other, only for data, so a warp whose sub-block is touched by few splats runs ahead by up to STAGES-1 batches
// instead of idling at a block barrier after every batch (34 % of all stall samples in the barrier version,
// profiles/r01c).  When every pixel of the tile is saturated the producer stops gathering (early termination).
#ifndef RS_RASTER_MIN_CTAS
#define RS_RASTER_MIN_CTAS 1 // resident CTAs per SM the compiler must leave registers for (experiments: 6, 7)
#endif
template <int CDIM, bool VEC_COLORS>
__global__ void __launch_bounds__(RAST_THREADS, (CDIM <= 4) ? RS_RASTER_MIN_CTAS : 1)
rs_raster_fwd_kernel(const rs_raster_fwd_args a, const int ch_off, const int ch_cnt) {
    using Cfg = RastCfg<CDIM>;
    constexpr int CP = Cfg::CP;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(16) float rast_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ int done_warps;

    const uint32_t tiles_per_image = (uint32_t)(a.tile_width * a.tile_height);
    const uint32_t image_id = blockIdx.x / tiles_per_image;
    const uint32_t tile_id = blockIdx.x - image_id * tiles_per_image;
    const uint32_t tile_y = tile_id / (uint32_t)a.tile_width;
    const uint32_t tile_x = tile_id - tile_y * (uint32_t)a.tile_width;

    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const bool producer = warp == RAST_CONSUMERS;
    // consumer warp -> 8x4 pixel sub-block, lane -> pixel inside it
    const uint32_t sub_x = tile_x * RS_TILE + (warp & 1) * 8;
    const uint32_t sub_y = tile_y * RS_TILE + ((warp >> 1) & 3) * 4;
    const uint32_t j = sub_x + (lane & 7);
    const uint32_t i = sub_y + (lane >> 3);
    const float px = (float)j + 0.5f;
    const float py = (float)i + 0.5f;
    const bool inside = !producer && (i < (uint32_t)a.image_height && j < (uint32_t)a.image_width);
    const size_t pix_id = (size_t)image_id * a.image_height * a.image_width + (size_t)i * a.image_width + j;

    const float *bg = a.backgrounds != nullptr ? a.backgrounds + (size_t)image_id * a.channels + ch_off : nullptr;

    // RasterizeToPixels3DGSFwd.cu:73-80: masked-out tile -> background colour only, alphas / last_ids untouched
    if (a.masks != nullptr && !a.masks[(size_t)image_id * tiles_per_image + tile_id]) {
        if (inside) {
            for (int k = 0; k < ch_cnt; ++k)
                a.render_colors[pix_id * a.channels + ch_off + k] = bg == nullptr ? 0.0f : bg[k];
            if (a.render_rgb8 != nullptr && ch_off == 0)
                for (int k = 0; k < 3 && k < ch_cnt; ++k) {
                    const float v = bg == nullptr ? 0.0f : bg[k];
                    a.render_rgb8[pix_id * 3 + k] =
                        (uint8_t)__float2uint_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.f), 0.5f), 0.f), 255.f));
                }
        }
        return;
    }

    const int64_t n_isects = a.n_isects_dev != nullptr ? min((int64_t)*a.n_isects_dev, a.n_isects) : a.n_isects;
    const int32_t *offs = a.tile_offsets + (size_t)image_id * tiles_per_image;
    const int32_t range_start = offs[tile_id];
    const int32_t range_end = (image_id == (uint32_t)a.I - 1 && tile_id == tiles_per_image - 1)
                                  ? (int32_t)n_isects
                                  : offs[tile_id + 1];
    const int num_batches = (range_end - range_start + RAST_BATCH - 1) / RAST_BATCH;

    if (tr == 0) {
        for (int s = 0; s < STAGES; ++s) {
            rs_mbar_init(&full_bar[s], 32);              // the 32 producer lanes
            rs_mbar_init(&empty_bar[s], RAST_CONSUMERS); // one arrival per consumer warp
        }
        done_warps = 0;
    }
    // pad columns of the colour rows are never copied; zero them once so the unpredicated FMAs below add 0
    if (!VEC_COLORS) {
        for (int r = tr; r < STAGES * RAST_BATCH; r += RAST_THREADS) {
            float *col = rast_smem + (size_t)(r / RAST_BATCH) * Cfg::STAGE_FLOATS + RAST_BATCH * 8 + (r % RAST_BATCH) * CP;
#pragma unroll
            for (int k = 0; k < CP; ++k)
                if (k >= ch_cnt)
                    col[k] = 0.f;
        }
    }
    __syncthreads(); // barriers initialised; the only block-wide barrier of the kernel
    volatile int *v_done = &done_warps;

    if (producer) {
        // ---------------------------------------------------------------------------------------------------------------
        // producer warp: lane l stages splats l, l+32, ..., l+224 of every batch
        // ---------------------------------------------------------------------------------------------------------------
        const float4 *records = reinterpret_cast<const float4 *>(a.records);
        constexpr int PER_LANE = RAST_BATCH / 32;
        // flatten ids are fetched TWO batches ahead of the copies that need them (gid = batch b, gnx = batch b+1), so the
        // id round trip is off the producer's critical path
        int32_t gid[PER_LANE], gnx[PER_LANE];
#pragma unroll
        for (int k = 0; k < PER_LANE; ++k) {
            const int32_t idx = range_start + k * 32 + lane;
            gid[k] = (idx < range_end) ? a.flatten_ids[idx] : -1;
            gnx[k] = (idx + RAST_BATCH < range_end) ? a.flatten_ids[idx + RAST_BATCH] : -1;
        }
        for (int b = 0; b < num_batches; ++b) {
            const int st = b % STAGES;
            const unsigned ph = (unsigned)(b / STAGES) & 1u;
            bool stop = false;
            while (!rs_mbar_try_wait(&empty_bar[st], ph ^ 1u)) { // stage released by all consumers (free at first use)
                if (*v_done >= RAST_CONSUMERS) {
                    stop = true;
                    break;
                }
            }
            if (stop || *v_done >= RAST_CONSUMERS)
                break; // every pixel saturated: the rest of the list is never read
            float *base = rast_smem + (size_t)st * Cfg::STAGE_FLOATS;
#pragma unroll
            for (int k = 0; k < PER_LANE; ++k) {
                const int32_t g = gid[k];
                if (g >= 0) {
                    const int t = k * 32 + lane;
                    float4 *r0 = reinterpret_cast<float4 *>(base) + t;
                    float4 *r1 = reinterpret_cast<float4 *>(base + RAST_BATCH * 4) + t;
                    float *col = base + RAST_BATCH * 8 + t * CP;
                    rs_cp_async16(r0, records + (size_t)g * 2);
                    rs_cp_async16(r1, records + (size_t)g * 2 + 1);
                    const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
                    const float *cp = a.colors + (size_t)gc * a.channels + ch_off;
                    if (VEC_COLORS) { // rows are 16-byte aligned and ch_cnt == CP
#pragma unroll
                        for (int c = 0; c < CP; c += 4)
                            rs_cp_async16(col + c, cp + c);
                    } else {
#pragma unroll
                        for (int c = 0; c < CDIM; ++c)
                            if (c < ch_cnt)
                                rs_cp_async4(col + c, cp + c);
                    }
                }
            }
            rs_cp_async_mbar_arrive(&full_bar[st]);
#pragma unroll
            for (int k = 0; k < PER_LANE; ++k) {
                gid[k] = gnx[k];
                const int32_t idx = range_start + RAST_BATCH * (b + 2) + k * 32 + lane;
                gnx[k] = (idx < range_end) ? a.flatten_ids[idx] : -1;
            }
        }
        rs_cp_async_wait_all(); // nothing may still be landing in shared memory when the CTA retires
        return;
    }

    // -------------------------------------------------------------------------------------------------------------------
    // consumer warps
    // -------------------------------------------------------------------------------------------------------------------
    // sub-block bounds in pixel-centre coordinates
    const float bx0 = (float)sub_x + 0.5f, bx1 = (float)sub_x + 7.5f;
    const float by0 = (float)sub_y + 0.5f, by1 = (float)sub_y + 3.5f;

    const unsigned smem_base = rs_smem_addr(rast_smem);
    float T = 1.0f;
    uint32_t cur_idx = 0;
    int done = inside ? 0 : 1; // int, not bool: the compiler keeps bools byte-packed (PRMT traffic in the loop)
    bool warp_done = __all_sync(0xffffffffu, done != 0);
    if (warp_done && lane == 0)
        atomicAdd(&done_warps, 1);
    float pix_out[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k)
        pix_out[k] = 0.f;

    for (int b = 0; b < num_batches; ++b) {
        const int st = b % STAGES;
        const unsigned ph = (unsigned)(b / STAGES) & 1u;
        bool stop = false;
        while (!rs_mbar_try_wait(&full_bar[st], ph)) {
            if (*v_done >= RAST_CONSUMERS) { // the producer has stopped (or will): nothing more to wait for
                stop = true;
                break;
            }
        }
        if (stop)
            break;
        if (!warp_done) {
            const int32_t batch_start = range_start + RAST_BATCH * b;
            const int batch_size = min(RAST_BATCH, range_end - batch_start);
            unsigned a_r0 = smem_base + (unsigned)(st * Cfg::STAGE_FLOATS * 4); // shared addresses of this stage
            asm volatile("mov.u32 %0, %0;\n" : "+r"(a_r0)); // opaque: keep it in a register, do not rematerialise per use
            const unsigned a_r1 = a_r0 + RAST_BATCH * 16;
            const unsigned a_col = a_r0 + RAST_BATCH * 32;

            for (int chunk = 0; chunk < batch_size; chunk += 32) {
#ifdef RS_RASTER_STATS
                if (lane == 0)
                    atomicAdd(&rs_stats[4], 1ull);
#endif
                const int t = chunk + lane;
                bool hit = false;
                if (t < batch_size) {
                    const float4 g0 = rs_lds128(a_r0 + t * 16);
                    const float4 g1 = rs_lds128(a_r1 + t * 16);
                    hit = rs_splat_touches_rect(g0.x, g0.y, g0.w, g1.x, g1.y, g1.z, bx0, bx1, by0, by1);
                }
                // bit-reversed ballot: the next splat in list order is the highest set bit (one FLO per iteration)
                unsigned m = __brev(__ballot_sync(0xffffffffu, hit));
                while (m) {
                    const int lz = __clz(m);
                    m &= ~(0x80000000u >> lz);
                    const int tt = chunk + lz;
#ifdef RS_RASTER_STATS // instrumentation build only (tools/raster_stats.py): how selective is the cull test?
                    {
                        const float4 q0 = rs_lds128(a_r0 + tt * 16);
                        const float4 q1 = rs_lds128(a_r1 + tt * 16);
                        const float ddx = q0.x - px, ddy = q0.y - py;
                        const float sg = 0.5f * (q0.w * ddx * ddx + q1.y * ddy * ddy) + q1.x * ddx * ddy;
                        const float al = fminf(0.999f, q0.z * __expf(-sg));
                        const bool pass = !done && !(sg < 0.f || al < RS_ALPHA_THRESHOLD);
                        const unsigned pm = __ballot_sync(0xffffffffu, pass);
                        const unsigned am = __ballot_sync(0xffffffffu, !done);
                        if (lane == 0) {
                            atomicAdd(&rs_stats[0], 1ull);
                            atomicAdd(&rs_stats[1], pm ? 1ull : 0ull);
                            atomicAdd(&rs_stats[2], (unsigned long long)__popc(pm));
                            atomicAdd(&rs_stats[3], (unsigned long long)__popc(am));
                        }
                    }
#endif
                    if (!done) {
                        const float4 g0 = rs_lds128(a_r0 + tt * 16);
                        const float4 g1 = rs_lds128(a_r1 + tt * 16);
                        const float dx = __fsub_rn(g0.x, px);
                        const float dy = __fsub_rn(g0.y, py);
                        const float tc = __fmul_rn(__fmul_rn(g1.y, dy), dy);
                        const float s = __fmaf_rn(dx, __fmul_rn(g0.w, dx), tc);
                        const float sigma = __fmaf_rn(dy, __fmul_rn(g1.x, dx), __fmul_rn(s, 0.5f));
                        const float alpha = fminf(0.999f, __fmul_rn(g0.z, __expf(-sigma)));
                        if (!(sigma < 0.f || alpha < RS_ALPHA_THRESHOLD)) {
                            const float next_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                            if (next_T <= 1e-4f) {
                                done = 1;
                            } else {
                                const float vis = __fmul_rn(alpha, T);
                                const unsigned crow = a_col + tt * (CP * 4);
#pragma unroll
                                for (int k = 0; k < CP; k += 4) {
                                    const float4 c4 = rs_lds128(crow + k * 4);
                                    pix_out[k + 0] = __fmaf_rn(c4.x, vis, pix_out[k + 0]);
                                    pix_out[k + 1] = __fmaf_rn(c4.y, vis, pix_out[k + 1]);
                                    pix_out[k + 2] = __fmaf_rn(c4.z, vis, pix_out[k + 2]);
                                    pix_out[k + 3] = __fmaf_rn(c4.w, vis, pix_out[k + 3]);
                                }
                                cur_idx = (uint32_t)(batch_start + tt);
                                T = next_T;
                            }
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done != 0)) {
                    warp_done = true;
                    if (lane == 0)
                        atomicAdd(&done_warps, 1);
                    break;
                }
            }
        }
        __syncwarp(); // every lane is finished with this stage
        if (lane == 0)
            rs_mbar_arrive(&empty_bar[st]);
    }

    if (inside) {
        a.render_alphas[pix_id] = __fsub_rn(1.0f, T);
        float *out = a.render_colors + pix_id * a.channels + ch_off;
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < ch_cnt)
                out[k] = bg == nullptr ? pix_out[k] : __fmaf_rn(T, bg[k], pix_out[k]);
        if (a.render_rgb8 != nullptr && ch_off == 0) { // 8-bit copy of the frame, quantised like torchvision save_image
            uint8_t *q = a.render_rgb8 + pix_id * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < CDIM) {
                    const float v = bg == nullptr ? pix_out[k] : __fmaf_rn(T, bg[k], pix_out[k]);
                    q[k] = (uint8_t)__float2uint_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.f), 0.5f), 0.f), 255.f));
                }
        }
        a.last_ids[pix_id] = (int32_t)cur_idx;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and the shared-memory ring runs ACROSS tile
// boundaries, so the producer warp fetches the first batches of the next tile (offsets -> flatten ids -> records: three
// dependent round trips) while the consumers are still compositing the current one, and no CTA is launched / retired per
// tile.  Every ring stage carries a small header {tile sequence number, tile, first intersection, size}; a consumer warp that
// sees a new sequence number writes out the pixel it has been accumulating and starts the next tile; a header with tile = -1
// ends the walk.  Early termination: when all eight consumer warps have saturated their pixels of tile s, the producer skips
// the rest of its list (done[s & 3]; the producer can be at most STAGES <= 3 headers ahead of the slowest consumer, so four
// counters never alias).  The per-pixel arithmetic and its order are those of the kernel above: results are bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
struct RastStageHdr {
    int seq, tile, start, size; // tile < 0: terminal; size == 0: empty or masked tile (flag in `start`: -2 = masked)
};

template <int CDIM, bool VEC_COLORS>
__global__ void __launch_bounds__(RAST_THREADS)
rs_raster_fwd_persistent_kernel(const rs_raster_fwd_args a, const int ch_off, const int ch_cnt) {
    using Cfg = RastCfg<CDIM>;
    constexpr int CP = Cfg::CP;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(16) float rast_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ RastStageHdr hdr[STAGES];
    __shared__ int done[4];

    const uint32_t tiles_per_image = (uint32_t)(a.tile_width * a.tile_height);
    const int n_tiles_total = (int)(tiles_per_image * (uint32_t)a.I);
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const bool producer = warp == RAST_CONSUMERS;
    const int64_t n_isects = a.n_isects_dev != nullptr ? min((int64_t)*a.n_isects_dev, a.n_isects) : a.n_isects;

    if (tr == 0) {
        for (int s = 0; s < STAGES; ++s) {
            rs_mbar_init(&full_bar[s], 33);              // 32 cp.async completions + the header writer's release-arrive
            rs_mbar_init(&empty_bar[s], RAST_CONSUMERS); // one arrival per consumer warp
        }
        done[0] = done[1] = done[2] = done[3] = 0;
    }
    if (!VEC_COLORS) {
        for (int r = tr; r < STAGES * RAST_BATCH; r += RAST_THREADS) {
            float *col = rast_smem + (size_t)(r / RAST_BATCH) * Cfg::STAGE_FLOATS + RAST_BATCH * 8 + (r % RAST_BATCH) * CP;
#pragma unroll
            for (int k = 0; k < CP; ++k)
                if (k >= ch_cnt)
                    col[k] = 0.f;
        }
    }
    __syncthreads();
    volatile int *v_done = done;

    if (producer) {
        const float4 *records = reinterpret_cast<const float4 *>(a.records);
        constexpr int PER_LANE = RAST_BATCH / 32;
        int gb = 0; // ring messages issued so far
        int seq = 0;
        auto send = [&](int tile, int start, int size) { // header of the message in stage gb % STAGES, then the arrive
            const int st = gb % STAGES;
            if (lane == 0) {
                hdr[st].seq = seq;
                hdr[st].tile = tile;
                hdr[st].start = start;
                hdr[st].size = size;
                rs_mbar_arrive(&full_bar[st]); // release: the header is visible to whoever observes the phase flip
            }
            rs_cp_async_mbar_arrive(&full_bar[st]);
            ++gb;
        };
        auto wait_free = [&]() {
            const int st = gb % STAGES;
            const unsigned ph = (unsigned)(gb / STAGES) & 1u;
            while (!rs_mbar_try_wait(&empty_bar[st], ph ^ 1u)) {
            }
        };
        // tiles: the first one is blockIdx.x; the following ones come from the work counter when the caller provides one
        // (dynamic: CTAs that drew short lists take more tiles), else in a fixed stride
        for (int tile = blockIdx.x; tile < n_tiles_total; ++seq) {
            if (lane == 0)
                v_done[seq & 3] = 0;
            __syncwarp();
            const int this_tile = tile;
            if (a.tile_counter != nullptr) {
                int nxt = 0;
                if (lane == 0)
                    nxt = (int)gridDim.x + (int)atomicAdd(a.tile_counter, 1u);
                tile = __shfl_sync(0xffffffffu, nxt, 0);
            } else {
                tile += gridDim.x;
            }
            const uint32_t image_id = (uint32_t)this_tile / tiles_per_image;
            const uint32_t tile_id = (uint32_t)this_tile - image_id * tiles_per_image;
            if (a.masks != nullptr && !a.masks[this_tile]) {
                wait_free();
                send(this_tile, -2, 0);
                continue;
            }
            const int32_t range_start = a.tile_offsets[this_tile];
            const int32_t range_end = (image_id == (uint32_t)a.I - 1 && tile_id == tiles_per_image - 1) ? (int32_t)n_isects
                                                                                                        : a.tile_offsets[this_tile + 1];
            const int num_batches = (range_end - range_start + RAST_BATCH - 1) / RAST_BATCH;
            if (num_batches <= 0) {
                wait_free();
                send(this_tile, range_start, 0);
                continue;
            }
            int32_t gid[PER_LANE], gnx[PER_LANE];
#pragma unroll
            for (int k = 0; k < PER_LANE; ++k) {
                const int32_t idx = range_start + k * 32 + lane;
                gid[k] = (idx < range_end) ? a.flatten_ids[idx] : -1;
                gnx[k] = (idx + RAST_BATCH < range_end) ? a.flatten_ids[idx + RAST_BATCH] : -1;
            }
            for (int b = 0; b < num_batches; ++b) {
                wait_free();
                if (v_done[seq & 3] >= RAST_CONSUMERS)
                    break; // every pixel of this tile saturated: the rest of its list is never read
                const int st = gb % STAGES;
                float *base = rast_smem + (size_t)st * Cfg::STAGE_FLOATS;
#pragma unroll
                for (int k = 0; k < PER_LANE; ++k) {
                    const int32_t g = gid[k];
                    if (g >= 0) {
                        const int t = k * 32 + lane;
                        float4 *r0 = reinterpret_cast<float4 *>(base) + t;
                        float4 *r1 = reinterpret_cast<float4 *>(base + RAST_BATCH * 4) + t;
                        float *col = base + RAST_BATCH * 8 + t * CP;
                        rs_cp_async16(r0, records + (size_t)g * 2);
                        rs_cp_async16(r1, records + (size_t)g * 2 + 1);
                        const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
                        const float *cp = a.colors + (size_t)gc * a.channels + ch_off;
                        if (VEC_COLORS) {
#pragma unroll
                            for (int c = 0; c < CP; c += 4)
                                rs_cp_async16(col + c, cp + c);
                        } else {
#pragma unroll
                            for (int c = 0; c < CDIM; ++c)
                                if (c < ch_cnt)
                                    rs_cp_async4(col + c, cp + c);
                        }
                    }
                }
                const int32_t batch_start = range_start + RAST_BATCH * b;
                send(this_tile, batch_start, min(RAST_BATCH, range_end - batch_start));
#pragma unroll
                for (int k = 0; k < PER_LANE; ++k) {
                    gid[k] = gnx[k];
                    const int32_t idx = range_start + RAST_BATCH * (b + 2) + k * 32 + lane;
                    gnx[k] = (idx < range_end) ? a.flatten_ids[idx] : -1;
                }
            }
        }
        wait_free();
        send(-1, 0, 0); // terminal
        rs_cp_async_wait_all();
        return;
    }

    // ---- consumer warps ------------------------------------------------------------------------------------------------
    const unsigned smem_base = rs_smem_addr(rast_smem);
    int cur_seq = -1;
    uint32_t image_id = 0;
    bool inside = false, masked = false;
    size_t pix_id = 0;
    float px = 0.f, py = 0.f, bx0 = 0.f, bx1 = 0.f, by0 = 0.f, by1 = 0.f;
    float T = 1.0f;
    uint32_t cur_idx = 0;
    int pdone = 1;
    bool warp_done = true;
    float pix_out[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k)
        pix_out[k] = 0.f;

    auto finalize = [&]() { // write the pixel of the tile that just ended
        if (cur_seq < 0 || !inside)
            return;
        const float *bg = a.backgrounds != nullptr ? a.backgrounds + (size_t)image_id * a.channels + ch_off : nullptr;
        float *out = a.render_colors + pix_id * a.channels + ch_off;
        if (masked) { // RasterizeToPixels3DGSFwd.cu:73-80: background colour only, alphas / last_ids untouched
            for (int k = 0; k < ch_cnt; ++k)
                out[k] = bg == nullptr ? 0.0f : bg[k];
            if (a.render_rgb8 != nullptr && ch_off == 0)
                for (int k = 0; k < 3 && k < ch_cnt; ++k) {
                    const float v = bg == nullptr ? 0.0f : bg[k];
                    a.render_rgb8[pix_id * 3 + k] =
                        (uint8_t)__float2uint_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.f), 0.5f), 0.f), 255.f));
                }
            return;
        }
        a.render_alphas[pix_id] = __fsub_rn(1.0f, T);
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < ch_cnt)
                out[k] = bg == nullptr ? pix_out[k] : __fmaf_rn(T, bg[k], pix_out[k]);
        if (a.render_rgb8 != nullptr && ch_off == 0) {
            uint8_t *q = a.render_rgb8 + pix_id * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < CDIM) {
                    const float v = bg == nullptr ? pix_out[k] : __fmaf_rn(T, bg[k], pix_out[k]);
                    q[k] = (uint8_t)__float2uint_rz(fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.f), 0.5f), 0.f), 255.f));
                }
        }
        a.last_ids[pix_id] = (int32_t)cur_idx;
    };

    for (int gb = 0;; ++gb) {
        const int st = gb % STAGES;
        const unsigned ph = (unsigned)(gb / STAGES) & 1u;
        while (!rs_mbar_try_wait(&full_bar[st], ph)) {
        }
        const int h_seq = hdr[st].seq, h_tile = hdr[st].tile, h_start = hdr[st].start, h_size = hdr[st].size;
        if (h_tile < 0) {
            finalize();
            break;
        }
        if (h_seq != cur_seq) {
            finalize();
            cur_seq = h_seq;
            image_id = (uint32_t)h_tile / tiles_per_image;
            const uint32_t tile_id = (uint32_t)h_tile - image_id * tiles_per_image;
            const uint32_t tile_y = tile_id / (uint32_t)a.tile_width;
            const uint32_t tile_x = tile_id - tile_y * (uint32_t)a.tile_width;
            const uint32_t sub_x = tile_x * RS_TILE + (warp & 1) * 8;
            const uint32_t sub_y = tile_y * RS_TILE + ((warp >> 1) & 3) * 4;
            const uint32_t j = sub_x + (lane & 7);
            const uint32_t i = sub_y + (lane >> 3);
            px = (float)j + 0.5f;
            py = (float)i + 0.5f;
            inside = i < (uint32_t)a.image_height && j < (uint32_t)a.image_width;
            pix_id = (size_t)image_id * a.image_height * a.image_width + (size_t)i * a.image_width + j;
            bx0 = (float)sub_x + 0.5f, bx1 = (float)sub_x + 7.5f;
            by0 = (float)sub_y + 0.5f, by1 = (float)sub_y + 3.5f;
            masked = h_start == -2 && h_size == 0;
            T = 1.0f;
            cur_idx = 0;
            pdone = inside ? 0 : 1;
#pragma unroll
            for (int k = 0; k < CP; ++k)
                pix_out[k] = 0.f;
            warp_done = __all_sync(0xffffffffu, pdone != 0);
            if (warp_done && lane == 0)
                atomicAdd(&done[cur_seq & 3], 1);
        }
        if (!warp_done && h_size > 0) {
            const int32_t batch_start = h_start;
            const int batch_size = h_size;
            unsigned a_r0 = smem_base + (unsigned)(st * Cfg::STAGE_FLOATS * 4);
            asm volatile("mov.u32 %0, %0;\n" : "+r"(a_r0));
            const unsigned a_r1 = a_r0 + RAST_BATCH * 16;
            const unsigned a_col = a_r0 + RAST_BATCH * 32;
            for (int chunk = 0; chunk < batch_size; chunk += 32) {
                const int t = chunk + lane;
                bool hit = false;
                if (t < batch_size) {
                    const float4 g0 = rs_lds128(a_r0 + t * 16);
                    const float4 g1 = rs_lds128(a_r1 + t * 16);
                    hit = rs_splat_touches_rect(g0.x, g0.y, g0.w, g1.x, g1.y, g1.z, bx0, bx1, by0, by1);
                }
                unsigned m = __brev(__ballot_sync(0xffffffffu, hit));
                while (m) {
                    const int lz = __clz(m);
                    m &= ~(0x80000000u >> lz);
                    const int tt = chunk + lz;
                    if (!pdone) {
                        const float4 g0 = rs_lds128(a_r0 + tt * 16);
                        const float4 g1 = rs_lds128(a_r1 + tt * 16);
                        const float dx = __fsub_rn(g0.x, px);
                        const float dy = __fsub_rn(g0.y, py);
                        const float tc = __fmul_rn(__fmul_rn(g1.y, dy), dy);
                        const float s = __fmaf_rn(dx, __fmul_rn(g0.w, dx), tc);
                        const float sigma = __fmaf_rn(dy, __fmul_rn(g1.x, dx), __fmul_rn(s, 0.5f));
                        const float alpha = fminf(0.999f, __fmul_rn(g0.z, __expf(-sigma)));
                        if (!(sigma < 0.f || alpha < RS_ALPHA_THRESHOLD)) {
                            const float next_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                            if (next_T <= 1e-4f) {
                                pdone = 1;
                            } else {
                                const float vis = __fmul_rn(alpha, T);
                                const unsigned crow = a_col + tt * (CP * 4);
#pragma unroll
                                for (int k = 0; k < CP; k += 4) {
                                    const float4 c4 = rs_lds128(crow + k * 4);
                                    pix_out[k + 0] = __fmaf_rn(c4.x, vis, pix_out[k + 0]);
                                    pix_out[k + 1] = __fmaf_rn(c4.y, vis, pix_out[k + 1]);
                                    pix_out[k + 2] = __fmaf_rn(c4.z, vis, pix_out[k + 2]);
                                    pix_out[k + 3] = __fmaf_rn(c4.w, vis, pix_out[k + 3]);
                                }
                                cur_idx = (uint32_t)(batch_start + tt);
                                T = next_T;
                            }
                        }
                    }
                }
                if (__all_sync(0xffffffffu, pdone != 0)) {
                    warp_done = true;
                    if (lane == 0)
                        atomicAdd(&done[cur_seq & 3], 1);
                    break;
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            rs_mbar_arrive(&empty_bar[st]);
    }
}

// RS_RASTER_PERSISTENT=0|1 selects the kernel at run time (A/B measurements); default: see raster_persistent_default()
static bool raster_persistent_default() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("RS_RASTER_PERSISTENT");
        cached = e != nullptr ? (atoi(e) != 0) : RS_RASTER_PERSISTENT_DEFAULT;
    }
    return cached != 0;
}

template <int CDIM>
static int launch_raster_fwd(const rs_raster_fwd_args &a, int ch_off, int ch_cnt, cudaStream_t s) {
    using Cfg = RastCfg<CDIM>;
    const int64_t grid = (int64_t)a.I * a.tile_width * a.tile_height;
    // 16-byte colour copies need aligned rows that fill the shared-memory pitch exactly
    const bool vec = (ch_cnt == Cfg::CP) && (a.channels % 4 == 0) && (ch_off % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.colors) & 15) == 0);
    if (raster_persistent_default()) {
        static RsPerDevice pattr[2];
        static int per_sm_cache[2] = {0, 0}; // resident CTAs per SM (registers, shared memory, threads): the grid is ONE wave
        auto kern = vec ? rs_raster_fwd_persistent_kernel<CDIM, true> : rs_raster_fwd_persistent_kernel<CDIM, false>;
        if (!rs_dev_done(pattr[vec])) {
            RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
            int per_sm = 0;
            RS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RAST_THREADS, Cfg::SMEM));
            per_sm_cache[vec] = per_sm > 0 ? per_sm : 1;
            rs_dev_mark(pattr[vec]);
        }
        const int64_t pgrid = min(grid, (int64_t)rs_num_sms() * per_sm_cache[vec]);
        kern<<<(unsigned)pgrid, RAST_THREADS, Cfg::SMEM, s>>>(a, ch_off, ch_cnt);
        RS_LAUNCH_CHECK("rs_raster_fwd_kernel");
        return 0;
    }
    static RsPerDevice attr_done[2];
    if (vec) {
        if (!rs_dev_done(attr_done[1])) {
            RS_CUDA(cudaFuncSetAttribute(rs_raster_fwd_kernel<CDIM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM));
            rs_dev_mark(attr_done[1]);
        }
        rs_raster_fwd_kernel<CDIM, true><<<(unsigned)grid, RAST_THREADS, Cfg::SMEM, s>>>(a, ch_off, ch_cnt);
    } else {
        if (!rs_dev_done(attr_done[0])) {
            RS_CUDA(cudaFuncSetAttribute(rs_raster_fwd_kernel<CDIM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM));
            rs_dev_mark(attr_done[0]);
        }
        rs_raster_fwd_kernel<CDIM, false><<<(unsigned)grid, RAST_THREADS, Cfg::SMEM, s>>>(a, ch_off, ch_cnt);
    }
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
    if (ch_cnt <= 8)
        return launch_raster_fwd<8>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 16)
        return launch_raster_fwd<16>(a, ch_off, ch_cnt, s);
    if (ch_cnt <= 20)
        return launch_raster_fwd<20>(a, ch_off, ch_cnt, s);
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

int rs_raster_pack_records(const rs_raster_fwd_args &a, cudaStream_t s) {
    RS_CHECK(a.n_rows > 0 || a.n_isects == 0, "rs_raster pack: n_rows required to pack records");
    RS_CHECK(a.n_rows == 0 || (a.means2d && a.conics && a.opacities && a.records), "rs_raster pack: null pointer");
    if (a.n_rows > 0) {
        rs_raster_pack_kernel<<<rs_cdiv(a.n_rows, 256), 256, 0, s>>>(a, a.n_rows);
        RS_LAUNCH_CHECK("rs_raster_pack_kernel");
    }
    return 0;
}

extern "C" int rs_raster_fwd(const rs_raster_fwd_args *a, rs_stream_t stream) {
    if (int e = rs_check_raster_args(a, "rs_raster_fwd"))
        return e;
    if (a->I == 0)
        return 0;
    // no projected splat at all (zero packed rows: every Gaussian culled): the kernel only writes backgrounds, and the
    // per-splat arrays may legitimately be empty (NULL)
    const bool no_rows = a->n_isects == 0 && a->n_isects_dev == nullptr;
    RS_CHECK((a->colors || no_rows) && a->tile_offsets && a->render_colors && a->render_alphas && a->last_ids,
             "rs_raster_fwd: null pointer");
    RS_CHECK(a->n_isects == 0 || a->flatten_ids != nullptr, "rs_raster_fwd: null flatten_ids");
    RS_CHECK(a->render_rgb8 == nullptr || a->channels >= 3, "rs_raster_fwd: render_rgb8 needs at least 3 channels");
    RS_CHECK((a->records != nullptr || no_rows) && (reinterpret_cast<uintptr_t>(a->records) & 15) == 0,
             "rs_raster_fwd: records scratch ([rows, 8] float, 16-byte aligned) is required");
    if (!a->records_ready && !no_rows) {
        if (int e = rs_raster_pack_records(*a, (cudaStream_t)stream))
            return e;
    }
    rs_raster_fwd_args chunk_args = *a;
    for (int off = 0; off < a->channels; off += 32) {
        const int cnt = a->channels - off < 32 ? a->channels - off : 32;
        if (int e = rs_raster_fwd_chunk(chunk_args, off, cnt, (cudaStream_t)stream))
            return e;
        chunk_args.tile_counter = nullptr; // the counter is consumed by the first launch; further chunks walk in a fixed stride
    }
    return 0;
}
