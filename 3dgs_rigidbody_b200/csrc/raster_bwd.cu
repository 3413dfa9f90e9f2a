// rs_raster_bwd: VJP of the compositing, back to front.
// Replaces csrc/RasterizeToPixels3DGSBwd.cu:15-276 (host side csrc/Rasterization.cpp:117-228).
//
// Same tile / sub-block decomposition as the forward kernel (one CTA per 16x16 tile, a warp per 8x4 pixels, per-warp
// exact culling of splats that cannot reach the sub-block), plus two B200-oriented changes to the reduction that dominates
// the reference kernel:
//   * the reference reduces each of the (CDIM + 8) gradient components with its own 5-step cg::reduce and lets lane 0
//     issue (CDIM + 8) serial atomics.  Here the components are reduced together with a folding butterfly: at each of
//     the 5 steps a lane keeps one half of its values and trades the other half, so the whole vector costs ~NV shuffles
//     instead of 5*NV, and the totals end up spread over NV lanes which then issue ONE predicated red.global each.
//   * per-lane destination pointers (which output array / component a lane ends up owning) are computed once per thread.
//   * a CTA-level reduction through a shared-memory accumulator (one red.global per (tile, splat, component)) was built
//     and measured SLOWER -- 1.63 ms vs 1.03 ms at c3 (profiles/r02_raster_bwd_smem_reduce_experiment.txt): float atomicAdd
//     on shared memory is a compare-and-swap loop that spins when the eight warps of a tile hit the same words, while
//     red.global is a fire-and-forget L2 operation.  The variant was removed again.
// Per-pixel math follows RasterizeToPixels3DGSBwd.cu:160-242.
#include "raster_common.cuh"

#define RAST_THREADS 256

int rs_check_raster_args(const rs_raster_fwd_args *a, const char *who);

template <int CDIM> struct RastBwdSmem {
    float4 xyoa[RAST_THREADS];
    float4 bcee[RAST_THREADS]; // conic b, conic c, cull limit (common.cuh: rs_cull_limit), unused
    int32_t id[RAST_THREADS];
    float color[CDIM][RAST_THREADS];
};

// Folding butterfly: on return, lane L holds in v[0] the warp total of component `comp(L)` (see fold_owner()).
template <int NV> __device__ __forceinline__ void warp_fold_reduce(float (&v)[NV], const int lane) {
    int n = NV;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int h = (n + 1) >> 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < (NV + 1) / 2; ++i) {
            if (i < h) {
                const float lo = v[i];
                const float hi = (h + i < n) ? v[(h + i < NV) ? h + i : 0] : 0.f;
                const float keep = upper ? hi : lo;
                const float send = upper ? lo : hi;
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        n = h;
    }
}
// number of values a lane is left with after the 5 folding steps (1 for NV <= 32, 2 for NV <= 64)
constexpr int fold_final_n(int nv) {
    int n = nv;
    for (int s = 0; s < 5; ++s)
        n = (n + 1) >> 1;
    return n;
}
// After warp_fold_reduce<NV>, lane `lane` holds components [base, base + real) in v[0..real); returns base, sets real.
template <int NV> __device__ __forceinline__ int fold_owner(const int lane, int &real) {
    int n = NV, base = 0, r = NV;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int h = (n + 1) >> 1;
        if (lane & o) {
            base += h;
            r = max(r - h, 0);
        } else {
            r = min(r, h);
        }
        n = h;
    }
    real = r;
    return base;
}

// component layout of the reduced vector: [0,CDIM) v_colors, CDIM..+2 v_conics, +3..+4 v_means2d, +5 v_opacity,
// +6..+7 v_means2d_abs (ABS only)
#ifndef RS_BWD_PREFETCH
#define RS_BWD_PREFETCH 1
#endif
template <int CDIM, bool ABS>
__global__ void __launch_bounds__(RAST_THREADS, (CDIM <= 16) ? 3 : 2)
rs_raster_bwd_kernel(const rs_raster_bwd_args b, const int ch_off, const int ch_cnt, const bool first_chunk) {
    constexpr int NV = CDIM + 6 + (ABS ? 2 : 0);
    __shared__ RastBwdSmem<CDIM> sm;
    const rs_raster_fwd_args &a = b.f;

    const uint32_t tiles_per_image = (uint32_t)(a.tile_width * a.tile_height);
    const uint32_t image_id = blockIdx.x / tiles_per_image;
    const uint32_t tile_id = blockIdx.x - image_id * tiles_per_image;
    const uint32_t tile_y = tile_id / (uint32_t)a.tile_width;
    const uint32_t tile_x = tile_id - tile_y * (uint32_t)a.tile_width;
    if (a.masks != nullptr && !a.masks[(size_t)image_id * tiles_per_image + tile_id])
        return;

    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const uint32_t sub_x = tile_x * RS_TILE + (warp & 1) * 8;
    const uint32_t sub_y = tile_y * RS_TILE + (warp >> 1) * 4;
    const uint32_t j = sub_x + (lane & 7);
    const uint32_t i = sub_y + (lane >> 3);
    const float px = (float)j + 0.5f;
    const float py = (float)i + 0.5f;
    const bool inside = (i < (uint32_t)a.image_height && j < (uint32_t)a.image_width);
    const size_t img_pix = (size_t)image_id * a.image_height * a.image_width;
    // clamp to the last pixel like the reference (Bwd.cu:78-79) so out-of-image lanes read valid memory
    const size_t pix_id =
        img_pix + min((size_t)i * a.image_width + j, (size_t)a.image_width * a.image_height - 1);

    const int64_t n_isects = a.n_isects_dev != nullptr ? min((int64_t)*a.n_isects_dev, a.n_isects) : a.n_isects;
    const int32_t *offs = a.tile_offsets + (size_t)image_id * tiles_per_image;
    const int32_t range_start = offs[tile_id];
    const int32_t range_end = (image_id == (uint32_t)a.I - 1 && tile_id == tiles_per_image - 1)
                                  ? (int32_t)n_isects
                                  : offs[tile_id + 1];
    const int num_batches = (range_end - range_start + RAST_THREADS - 1) / RAST_THREADS;

    const float bx0 = (float)sub_x + 0.5f, bx1 = (float)sub_x + 7.5f;
    const float by0 = (float)sub_y + 0.5f, by1 = (float)sub_y + 3.5f;

    const float T_final = 1.0f - a.render_alphas[pix_id];
    float T = T_final;
    float buffer[CDIM];
    float v_render_c[CDIM];
#pragma unroll
    for (int k = 0; k < CDIM; ++k) {
        buffer[k] = 0.f;
        v_render_c[k] = (k < ch_cnt) ? b.v_render_colors[pix_id * a.channels + ch_off + k] : 0.f;
    }
    // the alpha gradient enters once: with the first channel chunk only
    const float v_render_a = first_chunk ? b.v_render_alphas[pix_id] : 0.f;
    const int32_t bin_final = inside ? a.last_ids[pix_id] : 0;
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));
    const bool inside_any = __any_sync(0xffffffffu, inside);

    float bg_dot = 0.f; // sum_k bg_k * v_render_c_k  (Bwd.cu:210-217)
    if (a.backgrounds != nullptr) {
        const float *bg = a.backgrounds + (size_t)image_id * a.channels + ch_off;
#pragma unroll
        for (int k = 0; k < CDIM; ++k)
            if (k < ch_cnt)
                bg_dot += bg[k] * v_render_c[k];
    }

    // which reduced components this lane will own, and where they go
    constexpr int OWN = fold_final_n(NV);
    int own_real;
    const int own0 = fold_owner<NV>(lane, own_real);
    float *own_base[OWN];
    int own_stride[OWN];
    int own_kind[OWN]; // 0 = geometry (indexed by flatten id), 1 = colour row, 2 = opacity row
#pragma unroll
    for (int q = 0; q < OWN; ++q) {
        const int own = own0 + q;
        own_base[q] = nullptr;
        own_stride[q] = 0;
        own_kind[q] = 0;
        if (q < own_real) {
            if (own < CDIM) {
                if (own < ch_cnt) {
                    own_base[q] = b.v_colors + ch_off + own;
                    own_stride[q] = a.channels;
                    own_kind[q] = 1;
                }
            } else if (own < CDIM + 3) {
                own_base[q] = b.v_conics + (own - CDIM);
                own_stride[q] = 3;
            } else if (own < CDIM + 5) {
                own_base[q] = b.v_means2d + (own - CDIM - 3);
                own_stride[q] = 2;
            } else if (own < CDIM + 6) {
                own_base[q] = b.v_opacities;
                own_stride[q] = 1;
                own_kind[q] = 2;
            } else if (ABS) {
                own_base[q] = b.v_means2d_abs + (own - CDIM - 6);
                own_stride[q] = 2;
            }
        }
    }

    // Splats behind the furthest-back contributor of the whole TILE were never blended by any of its pixels (occluded):
    // their batches are not even loaded.  In dense scenes that is most of the list (the forward pass reads 14-18 % of it).
    __shared__ int tile_bin_final;
    if (tr == 0)
        tile_bin_final = range_start - 1;
    __syncthreads();
    if (lane == 0 && inside_any)
        atomicMax(&tile_bin_final, warp_bin_final);
    __syncthreads();
    const int first_batch = max(0, (range_end - 1 - tile_bin_final) / RAST_THREADS);


    // The id of this thread's splat in the NEXT batch is fetched one batch ahead and its rows are requested into L2, so a
    // batch change costs one exposed L2 round trip instead of two dependent DRAM round trips (flatten id -> attributes).
    auto fetch_id = [&](int bb) -> int32_t {
        const int32_t idx = range_end - 1 - RAST_THREADS * bb - tr;
        if (bb >= num_batches || idx < range_start)
            return -1;
        const int32_t g = a.flatten_ids[idx];
#if RS_BWD_PREFETCH
        const int32_t go = a.attr_mod_opacities > 0 ? g % a.attr_mod_opacities : g;
        const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.means2d + (size_t)g * 2));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.conics + (size_t)g * 3));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.opacities + go));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.colors + (size_t)gc * a.channels + ch_off));
#endif
        return g;
    };
    int32_t g_next = fetch_id(first_batch);
    for (int bb = first_batch; bb < num_batches; ++bb) {
        __syncthreads(); // every warp is finished with the previous batch
        // slot 0 of a batch is its furthest-back splat (Bwd.cu:132-150)
        const int32_t batch_end = range_end - 1 - RAST_THREADS * bb;
        const int32_t batch_size = min(RAST_THREADS, batch_end + 1 - range_start);
        const int32_t idx = batch_end - tr;
        const int32_t g_cur = g_next;
        if (idx >= range_start) {
            const int32_t g = g_cur;
            const int32_t go = a.attr_mod_opacities > 0 ? g % a.attr_mod_opacities : g;
            const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
            const float2 xy = reinterpret_cast<const float2 *>(a.means2d)[g];
            const float op = a.opacities[go];
            const float ca = a.conics[(size_t)g * 3 + 0];
            const float cb = a.conics[(size_t)g * 3 + 1];
            const float cc = a.conics[(size_t)g * 3 + 2];
            sm.xyoa[tr] = make_float4(xy.x, xy.y, op, ca);
            sm.bcee[tr] = make_float4(cb, cc, rs_cull_limit(ca, cb, cc, op), 0.f);
            sm.id[tr] = g;
            const float *cp = a.colors + (size_t)gc * a.channels + ch_off;
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                if (k < ch_cnt)
                    sm.color[k][tr] = cp[k];
        }
        g_next = fetch_id(bb + 1);
        __syncthreads();

        const int t_begin = max(0, batch_end - warp_bin_final);
        for (int chunk = (t_begin & ~31); chunk < batch_size; chunk += 32) {
            const int t = chunk + lane;
            bool hit = false;
            if (t >= t_begin && t < batch_size) {
                const float4 g0 = sm.xyoa[t];
                const float4 g1 = sm.bcee[t];
                hit = rs_splat_touches_rect(g0.x, g0.y, g0.w, g1.x, g1.y, g1.z, bx0, bx1, by0, by1);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int tt = chunk + __ffs(m) - 1;
                m &= m - 1;
                bool valid = inside && (batch_end - tt <= bin_final);
                float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
                float ca = 0.f, cb = 0.f, cc = 0.f;
                if (valid) {
                    const float4 g0 = sm.xyoa[tt];
                    const float4 g1 = sm.bcee[tt];
                    opac = g0.z;
                    ca = g0.w;
                    cb = g1.x;
                    cc = g1.y;
                    dx = __fsub_rn(g0.x, px);
                    dy = __fsub_rn(g0.y, py);
                    const float tc = __fmul_rn(__fmul_rn(cc, dy), dy);
                    const float s = __fmaf_rn(dx, __fmul_rn(ca, dx), tc);
                    const float sigma = __fmaf_rn(dy, __fmul_rn(cb, dx), __fmul_rn(s, 0.5f));
                    vis = __expf(-sigma);
                    alpha = fminf(0.999f, __fmul_rn(opac, vis));
                    if (sigma < 0.f || alpha < RS_ALPHA_THRESHOLD)
                        valid = false;
                }
                if (!__any_sync(0xffffffffu, valid))
                    continue;

                float v[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    v[k] = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.f;
#pragma unroll
                    for (int k = 0; k < CDIM; ++k) {
                        if (k < ch_cnt) {
                            const float c = sm.color[k][tt];
                            v[k] = fac * v_render_c[k];
                            v_alpha += (c * T - buffer[k] * ra) * v_render_c[k];
                            buffer[k] += c * fac;
                        }
                    }
                    v_alpha += T_final * ra * v_render_a;
                    if (a.backgrounds != nullptr)
                        v_alpha += -T_final * ra * bg_dot;
                    if (opac * vis <= 0.999f) {
                        const float v_sigma = -opac * vis * v_alpha;
                        v[CDIM + 0] = 0.5f * v_sigma * dx * dx;
                        v[CDIM + 1] = v_sigma * dx * dy;
                        v[CDIM + 2] = 0.5f * v_sigma * dy * dy;
                        const float vx = v_sigma * (ca * dx + cb * dy);
                        const float vy = v_sigma * (cb * dx + cc * dy);
                        v[CDIM + 3] = vx;
                        v[CDIM + 4] = vy;
                        v[CDIM + 5] = vis * v_alpha;
                        if (ABS) {
                            v[CDIM + 6] = fabsf(vx);
                            v[CDIM + 7] = fabsf(vy);
                        }
                    }
                }
                warp_fold_reduce<NV>(v, lane);
                const int32_t g = sm.id[tt];
#pragma unroll
                for (int q = 0; q < OWN; ++q) {
                    if (own_base[q] != nullptr) {
                        int32_t row = g;
                        if (own_kind[q] == 1 && a.attr_mod_colors > 0)
                            row = g % a.attr_mod_colors;
                        else if (own_kind[q] == 2 && a.attr_mod_opacities > 0)
                            row = g % a.attr_mod_opacities;
                        atomicAdd(own_base[q] + (size_t)row * own_stride[q], v[q]);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Ring variant (RS_BWD_RING, default): the same per-pixel math and the same per-warp reduction, but the batches travel
// through a 3-stage shared-memory ring filled by a producer warp (warp 8) with cp.async completing on mbarriers, as in
// the forward kernel.  The eight compositing warps of a tile no longer meet at a __syncthreads per batch: ncu showed the
// barrier as the top stall reason of the kernel above (3.6 warps per issue slot), because the warps of a tile hit very
// different numbers of splats.  Batches of 128 splats, 32-byte records (packed by rs_raster_pack_records) + colour rows
// + flatten ids per stage.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RS_BWD_RING
#define RS_BWD_RING 1
#endif
#ifndef RS_BWD_RING_CTAS
#define RS_BWD_RING_CTAS 3
#endif
#define BWD_BATCH 128
#define BWD_STAGES 3
#define BWD_CONSUMERS 8
#define BWD_THREADS (32 * (BWD_CONSUMERS + 1))
template <int CDIM> struct BwdRingCfg {
    static constexpr int CP = (CDIM + 3) & ~3;
    static constexpr int STAGE_FLOATS = BWD_BATCH * (8 + CP + 1); // records (2 x float4), colour row, flatten id
    // Wide colour rows: the per-pixel cotangent v_render_colors[CDIM] lives in shared memory ([CP/4][256 pixels] float4,
    // conflict-free LDS.128, each lane reads only what it wrote) instead of CDIM registers -- with it, buffer[CDIM] and
    // the NV-wide reduction vector the 16-channel kernel spilled 124 bytes per thread at 72 registers, and ncu showed the
    // spill reloads as the top stall (long scoreboard 2.6 warps per issue slot).
    static constexpr bool VRC_SMEM = CDIM > 8;
    static constexpr int VRC_FLOATS = VRC_SMEM ? CP * 32 * BWD_CONSUMERS : 0;
    static constexpr size_t SMEM = ((size_t)BWD_STAGES * STAGE_FLOATS + VRC_FLOATS) * sizeof(float);
};

template <int CDIM, bool ABS>
__global__ void __launch_bounds__(BWD_THREADS, (CDIM <= 16) ? RS_BWD_RING_CTAS : 2)
rs_raster_bwd_ring_kernel(const rs_raster_bwd_args b, const int ch_off, const int ch_cnt, const bool first_chunk) {
    using Cfg = BwdRingCfg<CDIM>;
    constexpr int CP = Cfg::CP;
    constexpr int NV = CDIM + 6 + (ABS ? 2 : 0);
    extern __shared__ __align__(16) float ring[];
    __shared__ uint64_t full_bar[BWD_STAGES], empty_bar[BWD_STAGES];
    __shared__ int tile_bin_final;
    const rs_raster_fwd_args &a = b.f;

    const uint32_t tiles_per_image = (uint32_t)(a.tile_width * a.tile_height);
    const uint32_t image_id = blockIdx.x / tiles_per_image;
    const uint32_t tile_id = blockIdx.x - image_id * tiles_per_image;
    const uint32_t tile_y = tile_id / (uint32_t)a.tile_width;
    const uint32_t tile_x = tile_id - tile_y * (uint32_t)a.tile_width;
    if (a.masks != nullptr && !a.masks[(size_t)image_id * tiles_per_image + tile_id])
        return;

    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const bool producer = warp == BWD_CONSUMERS;
    const int cw = producer ? 0 : warp;
    const uint32_t sub_x = tile_x * RS_TILE + (cw & 1) * 8;
    const uint32_t sub_y = tile_y * RS_TILE + (cw >> 1) * 4;
    const uint32_t j = sub_x + (lane & 7);
    const uint32_t i = sub_y + (lane >> 3);
    const float px = (float)j + 0.5f;
    const float py = (float)i + 0.5f;
    const bool inside = !producer && (i < (uint32_t)a.image_height && j < (uint32_t)a.image_width);
    const size_t img_pix = (size_t)image_id * a.image_height * a.image_width;
    const size_t pix_id =
        img_pix + min((size_t)i * a.image_width + j, (size_t)a.image_width * a.image_height - 1);

    const int64_t n_isects = a.n_isects_dev != nullptr ? min((int64_t)*a.n_isects_dev, a.n_isects) : a.n_isects;
    const int32_t *offs = a.tile_offsets + (size_t)image_id * tiles_per_image;
    const int32_t range_start = offs[tile_id];
    const int32_t range_end = (image_id == (uint32_t)a.I - 1 && tile_id == tiles_per_image - 1)
                                  ? (int32_t)n_isects
                                  : offs[tile_id + 1];
    const int num_batches = (range_end - range_start + BWD_BATCH - 1) / BWD_BATCH;

    const int32_t bin_final = inside ? a.last_ids[pix_id] : 0;
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));
    const bool inside_any = __any_sync(0xffffffffu, inside);

    if (tr == 0) {
        for (int s = 0; s < BWD_STAGES; ++s) {
            rs_mbar_init(&full_bar[s], 32);             // the 32 producer lanes
            rs_mbar_init(&empty_bar[s], BWD_CONSUMERS); // one arrival per compositing warp
        }
        tile_bin_final = range_start - 1;
    }
    __syncthreads();
    if (lane == 0 && inside_any)
        atomicMax(&tile_bin_final, warp_bin_final);
    __syncthreads(); // the last block-wide barrier of the kernel
    // Splats behind the furthest-back contributor of the whole TILE were never blended by any of its pixels: their
    // batches are not even loaded.
    const int first_batch = max(0, (range_end - 1 - tile_bin_final) / BWD_BATCH);

    if (producer) {
        // slot t of batch bb holds list entry range_end - 1 - BWD_BATCH * bb - t (slot 0 = furthest back, Bwd.cu:132-150)
        const float4 *records = reinterpret_cast<const float4 *>(a.records);
        const bool vec = ch_cnt == CP && (a.channels & 3) == 0 && (ch_off & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(a.colors) & 15) == 0;
        constexpr int PER_LANE = BWD_BATCH / 32;
        int32_t gid[PER_LANE], gnx[PER_LANE]; // flatten ids of batch bb and bb + 1 (fetched two batches ahead)
#pragma unroll
        for (int k = 0; k < PER_LANE; ++k) {
            const int32_t idx = range_end - 1 - BWD_BATCH * first_batch - (k * 32 + lane);
            gid[k] = (first_batch < num_batches && idx >= range_start) ? a.flatten_ids[idx] : -1;
            gnx[k] = (first_batch + 1 < num_batches && idx - BWD_BATCH >= range_start) ? a.flatten_ids[idx - BWD_BATCH] : -1;
        }
        for (int bb = first_batch; bb < num_batches; ++bb) {
            const int it = bb - first_batch;
            const int st = it % BWD_STAGES;
            const unsigned ph = (unsigned)(it / BWD_STAGES) & 1u;
            while (!rs_mbar_try_wait(&empty_bar[st], ph ^ 1u)) { // released by all consumers (free at first use)
            }
            float *base = ring + (size_t)st * Cfg::STAGE_FLOATS;
#pragma unroll
            for (int k = 0; k < PER_LANE; ++k) {
                const int32_t g = gid[k];
                if (g >= 0) {
                    const int t = k * 32 + lane;
                    rs_cp_async16(reinterpret_cast<float4 *>(base) + t, records + (size_t)g * 2);
                    rs_cp_async16(reinterpret_cast<float4 *>(base + BWD_BATCH * 4) + t, records + (size_t)g * 2 + 1);
                    float *col = base + BWD_BATCH * 8 + t * CP;
                    const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
                    const float *cp = a.colors + (size_t)gc * a.channels + ch_off;
                    if (vec) {
#pragma unroll
                        for (int c = 0; c < CP; c += 4)
                            rs_cp_async16(col + c, cp + c);
                    } else {
#pragma unroll
                        for (int c = 0; c < CDIM; ++c)
                            if (c < ch_cnt)
                                rs_cp_async4(col + c, cp + c);
                    }
                    // the id travels the same way as the rows it names (no plain store to order against the barrier)
                    rs_cp_async4(base + BWD_BATCH * (8 + CP) + t, a.flatten_ids + (range_end - 1 - BWD_BATCH * bb - t));
                }
            }
            rs_cp_async_mbar_arrive(&full_bar[st]); // arrives once this lane's copies have landed
#pragma unroll
            for (int k = 0; k < PER_LANE; ++k) {
                gid[k] = gnx[k];
                const int32_t idx = range_end - 1 - BWD_BATCH * (bb + 2) - (k * 32 + lane);
                gnx[k] = (bb + 2 < num_batches && idx >= range_start) ? a.flatten_ids[idx] : -1;
            }
        }
        rs_cp_async_wait_all(); // nothing may still be landing in shared memory when the CTA retires
        return;
    }

    // -------------------------------------------------------------------------------------------------------------------
    // compositing warps
    // -------------------------------------------------------------------------------------------------------------------
    const float bx0 = (float)sub_x + 0.5f, bx1 = (float)sub_x + 7.5f;
    const float by0 = (float)sub_y + 0.5f, by1 = (float)sub_y + 3.5f;

    const float T_final = 1.0f - a.render_alphas[pix_id];
    float T = T_final;
    float buffer[CDIM];
    constexpr bool VRC_SMEM = Cfg::VRC_SMEM;
    float v_render_c[VRC_SMEM ? 1 : CDIM];
    const unsigned smem_base = rs_smem_addr(ring);
    // (VRC_SMEM) this pixel's cotangent: float4 number k4 at [k4][tr]; written and read by this lane only
    const unsigned a_vrc = smem_base + (unsigned)(BWD_STAGES * Cfg::STAGE_FLOATS * 4) + (unsigned)tr * 16u;
    float bg_dot = 0.f; // sum_k bg_k * v_render_c_k  (Bwd.cu:210-217)
    {
        const float *bg = a.backgrounds != nullptr ? a.backgrounds + (size_t)image_id * a.channels + ch_off : nullptr;
        const float *vr = b.v_render_colors + pix_id * a.channels + ch_off;
#pragma unroll
        for (int k4 = 0; k4 < CP; k4 += 4) {
            float vv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k4 + q;
                vv[q] = (k < CDIM && k < ch_cnt) ? vr[k] : 0.f;
                if (k < CDIM) {
                    buffer[k] = 0.f;
                    if (!VRC_SMEM)
                        v_render_c[VRC_SMEM ? 0 : k] = vv[q];
                    if (bg != nullptr && k < ch_cnt)
                        bg_dot += bg[k] * vv[q];
                }
            }
            if (VRC_SMEM)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a_vrc + (unsigned)(k4 / 4) * (BWD_CONSUMERS * 32 * 16)),
                             "f"(vv[0]), "f"(vv[1]), "f"(vv[2]), "f"(vv[3])
                             : "memory");
        }
    }
    const float v_render_a = first_chunk ? b.v_render_alphas[pix_id] : 0.f; // enters with the first channel chunk only

    // which reduced components this lane will own, and where they go
    constexpr int OWN = fold_final_n(NV);
    int own_real;
    const int own0 = fold_owner<NV>(lane, own_real);
    float *own_base[OWN];
    int own_stride[OWN];
    int own_kind[OWN]; // 0 = geometry (indexed by flatten id), 1 = colour row, 2 = opacity row
#pragma unroll
    for (int q = 0; q < OWN; ++q) {
        const int own = own0 + q;
        own_base[q] = nullptr;
        own_stride[q] = 0;
        own_kind[q] = 0;
        if (q < own_real) {
            if (own < CDIM) {
                if (own < ch_cnt) {
                    own_base[q] = b.v_colors + ch_off + own;
                    own_stride[q] = a.channels;
                    own_kind[q] = 1;
                }
            } else if (own < CDIM + 3) {
                own_base[q] = b.v_conics + (own - CDIM);
                own_stride[q] = 3;
            } else if (own < CDIM + 5) {
                own_base[q] = b.v_means2d + (own - CDIM - 3);
                own_stride[q] = 2;
            } else if (own < CDIM + 6) {
                own_base[q] = b.v_opacities;
                own_stride[q] = 1;
                own_kind[q] = 2;
            } else if (ABS) {
                own_base[q] = b.v_means2d_abs + (own - CDIM - 6);
                own_stride[q] = 2;
            }
        }
    }

    for (int bb = first_batch; bb < num_batches; ++bb) {
        const int it = bb - first_batch;
        const int st = it % BWD_STAGES;
        const unsigned ph = (unsigned)(it / BWD_STAGES) & 1u;
        while (!rs_mbar_try_wait(&full_bar[st], ph)) {
        }
        const int32_t batch_end = range_end - 1 - BWD_BATCH * bb;
        const int32_t batch_size = min(BWD_BATCH, batch_end + 1 - range_start);
        unsigned a_r0 = smem_base + (unsigned)(st * Cfg::STAGE_FLOATS * 4);
        asm volatile("mov.u32 %0, %0;\n" : "+r"(a_r0)); // opaque: keep it in a register
        const unsigned a_r1 = a_r0 + BWD_BATCH * 16;
        const unsigned a_col = a_r0 + BWD_BATCH * 32;
        const unsigned a_id = a_col + BWD_BATCH * CP * 4;

        const int t_begin = max(0, batch_end - warp_bin_final);
        for (int chunk = (t_begin & ~31); chunk < batch_size; chunk += 32) {
            const int t = chunk + lane;
            bool hit = false;
            if (t >= t_begin && t < batch_size) {
                const float4 g0 = rs_lds128(a_r0 + t * 16);
                const float4 g1 = rs_lds128(a_r1 + t * 16);
                hit = rs_splat_touches_rect(g0.x, g0.y, g0.w, g1.x, g1.y, g1.z, bx0, bx1, by0, by1);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int tt = chunk + __ffs(m) - 1;
                m &= m - 1;
                bool valid = inside && (batch_end - tt <= bin_final);
                float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
                float ca = 0.f, cb = 0.f, cc = 0.f;
                if (valid) {
                    const float4 g0 = rs_lds128(a_r0 + tt * 16);
                    const float4 g1 = rs_lds128(a_r1 + tt * 16);
                    opac = g0.z;
                    ca = g0.w;
                    cb = g1.x;
                    cc = g1.y;
                    dx = __fsub_rn(g0.x, px);
                    dy = __fsub_rn(g0.y, py);
                    const float tc = __fmul_rn(__fmul_rn(cc, dy), dy);
                    const float s = __fmaf_rn(dx, __fmul_rn(ca, dx), tc);
                    const float sigma = __fmaf_rn(dy, __fmul_rn(cb, dx), __fmul_rn(s, 0.5f));
                    vis = __expf(-sigma);
                    alpha = fminf(0.999f, __fmul_rn(opac, vis));
                    if (sigma < 0.f || alpha < RS_ALPHA_THRESHOLD)
                        valid = false;
                }
                if (!__any_sync(0xffffffffu, valid))
                    continue;

                float v[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    v[k] = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.f;
                    const unsigned crow = a_col + tt * (CP * 4);
#pragma unroll
                    for (int k4 = 0; k4 < CP; k4 += 4) {
                        const float4 c4 = rs_lds128(crow + k4 * 4);
                        const float cs[4] = {c4.x, c4.y, c4.z, c4.w};
                        float vs[4] = {0.f, 0.f, 0.f, 0.f};
                        if (VRC_SMEM) {
                            const float4 v4 = rs_lds128(a_vrc + (unsigned)(k4 / 4) * (BWD_CONSUMERS * 32 * 16));
                            vs[0] = v4.x, vs[1] = v4.y, vs[2] = v4.z, vs[3] = v4.w;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int k = k4 + q;
                            if (k < CDIM && k < ch_cnt) {
                                const float c = cs[q];
                                const float vr = VRC_SMEM ? vs[q] : v_render_c[VRC_SMEM ? 0 : k];
                                v[k] = fac * vr;
                                v_alpha += (c * T - buffer[k] * ra) * vr;
                                buffer[k] += c * fac;
                            }
                        }
                    }
                    v_alpha += T_final * ra * v_render_a;
                    if (a.backgrounds != nullptr)
                        v_alpha += -T_final * ra * bg_dot;
                    if (opac * vis <= 0.999f) {
                        const float v_sigma = -opac * vis * v_alpha;
                        v[CDIM + 0] = 0.5f * v_sigma * dx * dx;
                        v[CDIM + 1] = v_sigma * dx * dy;
                        v[CDIM + 2] = 0.5f * v_sigma * dy * dy;
                        const float vx = v_sigma * (ca * dx + cb * dy);
                        const float vy = v_sigma * (cb * dx + cc * dy);
                        v[CDIM + 3] = vx;
                        v[CDIM + 4] = vy;
                        v[CDIM + 5] = vis * v_alpha;
                        if (ABS) {
                            v[CDIM + 6] = fabsf(vx);
                            v[CDIM + 7] = fabsf(vy);
                        }
                    }
                }
                warp_fold_reduce<NV>(v, lane);
                const int32_t g = rs_lds32i(a_id + tt * 4);
#pragma unroll
                for (int q = 0; q < OWN; ++q) {
                    if (own_base[q] != nullptr) {
                        int32_t row = g;
                        if (own_kind[q] == 1 && a.attr_mod_colors > 0)
                            row = g % a.attr_mod_colors;
                        else if (own_kind[q] == 2 && a.attr_mod_opacities > 0)
                            row = g % a.attr_mod_opacities;
                        atomicAdd(own_base[q] + (size_t)row * own_stride[q], v[q]);
                    }
                }
            }
        }
        __syncwarp(); // every lane is finished with this stage
        if (lane == 0)
            rs_mbar_arrive(&empty_bar[st]);
    }
}

template <int CDIM>
static int launch_raster_bwd(const rs_raster_bwd_args &b, int ch_off, int ch_cnt, bool first, cudaStream_t s) {
    const int64_t grid = (int64_t)b.f.I * b.f.tile_width * b.f.tile_height;
    const bool abs_grad = b.v_means2d_abs != nullptr;
#if RS_BWD_RING
    if (b.f.records != nullptr) { // staged through the ring (records packed by rs_raster_bwd)
        static RsPerDevice ring_attr[2];
        const size_t ring_bytes = BwdRingCfg<CDIM>::SMEM;
        if (ring_bytes > 47 * 1024 && !rs_dev_done(ring_attr[abs_grad])) {
            if (abs_grad)
                RS_CUDA(cudaFuncSetAttribute(rs_raster_bwd_ring_kernel<CDIM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)ring_bytes));
            else
                RS_CUDA(cudaFuncSetAttribute(rs_raster_bwd_ring_kernel<CDIM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)ring_bytes));
            rs_dev_mark(ring_attr[abs_grad]);
        }
        if (abs_grad)
            rs_raster_bwd_ring_kernel<CDIM, true><<<(unsigned)grid, BWD_THREADS, ring_bytes, s>>>(b, ch_off, ch_cnt, first);
        else
            rs_raster_bwd_ring_kernel<CDIM, false><<<(unsigned)grid, BWD_THREADS, ring_bytes, s>>>(b, ch_off, ch_cnt, first);
        RS_LAUNCH_CHECK("rs_raster_bwd_ring_kernel");
        return 0;
    }
#endif
    if (abs_grad)
        rs_raster_bwd_kernel<CDIM, true><<<(unsigned)grid, RAST_THREADS, 0, s>>>(b, ch_off, ch_cnt, first);
    else
        rs_raster_bwd_kernel<CDIM, false><<<(unsigned)grid, RAST_THREADS, 0, s>>>(b, ch_off, ch_cnt, first);
    RS_LAUNCH_CHECK("rs_raster_bwd_kernel");
    return 0;
}

extern "C" int rs_raster_bwd(const rs_raster_bwd_args *b, rs_stream_t stream) {
    RS_CHECK(b != nullptr, "rs_raster_bwd: null args");
    if (int e = rs_check_raster_args(&b->f, "rs_raster_bwd"))
        return e;
    const rs_raster_fwd_args &a = b->f;
    const int64_t n = a.n_isects;
    // RasterizeToPixels3DGSBwd.cu:326-329: nothing to do without intersections
    if (a.I == 0 || (n == 0 && a.n_isects_dev == nullptr))
        return 0;
    RS_CHECK(a.means2d && a.conics && a.colors && a.opacities && a.tile_offsets && a.flatten_ids && a.render_alphas &&
                 a.last_ids && b->v_render_colors && b->v_render_alphas && b->v_means2d && b->v_conics &&
                 b->v_colors && b->v_opacities,
             "rs_raster_bwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (RS_BWD_RING && a.records != nullptr) {
        RS_CHECK((reinterpret_cast<uintptr_t>(a.records) & 15) == 0, "rs_raster_bwd: records must be 16-byte aligned");
        if (!a.records_ready)
            if (int e = rs_raster_pack_records(a, s))
                return e;
    }
    for (int off = 0; off < a.channels; off += 32) {
        const int cnt = a.channels - off < 32 ? a.channels - off : 32;
        int e;
        if (cnt <= 1)
            e = launch_raster_bwd<1>(*b, off, cnt, off == 0, s);
        else if (cnt <= 2)
            e = launch_raster_bwd<2>(*b, off, cnt, off == 0, s);
        else if (cnt <= 3)
            e = launch_raster_bwd<3>(*b, off, cnt, off == 0, s);
        else if (cnt <= 4)
            e = launch_raster_bwd<4>(*b, off, cnt, off == 0, s);
        else if (cnt <= 5)
            e = launch_raster_bwd<5>(*b, off, cnt, off == 0, s);
        else if (cnt <= 8)
            e = launch_raster_bwd<8>(*b, off, cnt, off == 0, s);
        else if (cnt <= 16)
            e = launch_raster_bwd<16>(*b, off, cnt, off == 0, s);
        else if (cnt <= 17)
            e = launch_raster_bwd<17>(*b, off, cnt, off == 0, s);
        else
            e = launch_raster_bwd<32>(*b, off, cnt, off == 0, s);
        if (e)
            return e;
    }
    return 0;
}
