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
//   * (experiment, off by default: -DRS_BWD_SMEM_REDUCE=1) CTA-level reduction: the owning lanes add into a shared-memory
//     accumulator [256 splats][NV] and the CTA flushes ONE red.global per (tile, splat, component) per batch.  Measured
//     SLOWER -- 1.63 ms vs 1.03 ms at c3 (profiles/r02_raster_bwd_smem_reduce_experiment.txt): float atomicAdd on shared
//     memory is a compare-and-swap loop that spins when the eight warps of a tile hit the same words, while red.global is
//     a fire-and-forget L2 operation.
// Per-pixel math follows RasterizeToPixels3DGSBwd.cu:160-242.
#include "common.cuh"

#define RAST_THREADS 256
#ifndef RS_BWD_SMEM_REDUCE
#define RS_BWD_SMEM_REDUCE 0
#endif

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
    constexpr int NVP = NV | 1; // odd pitch of the per-splat accumulator rows
    __shared__ RastBwdSmem<CDIM> sm;
    extern __shared__ __align__(16) float bwd_acc[]; // [RAST_THREADS][NVP] (RS_BWD_SMEM_REDUCE)
    __shared__ unsigned int touched[RAST_THREADS / 32];
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
#if RS_BWD_SMEM_REDUCE
#pragma unroll
    for (int k = 0; k < NVP; ++k)
        bwd_acc[tr * NVP + k] = 0.f;
    if (tr < RAST_THREADS / 32)
        touched[tr] = 0u;
#endif
    __syncthreads();
    if (lane == 0 && inside_any)
        atomicMax(&tile_bin_final, warp_bin_final);
    __syncthreads();
    const int first_batch = max(0, (range_end - 1 - tile_bin_final) / RAST_THREADS);

    // flush of the CTA-level accumulator: thread t owns splat slot t of the batch that was just processed
    auto flush = [&]() {
#if RS_BWD_SMEM_REDUCE
        if ((touched[tr >> 5] >> (tr & 31)) & 1u) {
            const int32_t g = sm.id[tr];
            const int32_t gc = a.attr_mod_colors > 0 ? g % a.attr_mod_colors : g;
            const int32_t go = a.attr_mod_opacities > 0 ? g % a.attr_mod_opacities : g;
            float *acc = bwd_acc + tr * NVP;
#pragma unroll
            for (int k = 0; k < CDIM; ++k)
                if (k < ch_cnt) {
                    atomicAdd(b.v_colors + (size_t)gc * a.channels + ch_off + k, acc[k]);
                    acc[k] = 0.f;
                }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                atomicAdd(b.v_conics + (size_t)g * 3 + k, acc[CDIM + k]);
                acc[CDIM + k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                atomicAdd(b.v_means2d + (size_t)g * 2 + k, acc[CDIM + 3 + k]);
                acc[CDIM + 3 + k] = 0.f;
                if (ABS) {
                    atomicAdd(b.v_means2d_abs + (size_t)g * 2 + k, acc[CDIM + 6 + k]);
                    acc[CDIM + 6 + k] = 0.f;
                }
            }
            atomicAdd(b.v_opacities + go, acc[CDIM + 5]);
            acc[CDIM + 5] = 0.f;
        }
        __syncwarp();
        if ((tr & 31) == 0)
            touched[tr >> 5] = 0u;
#endif
    };

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
        __syncthreads();
        if (bb > first_batch)
            flush(); // the previous batch (reads its sm.id[tr] before the load below replaces it)
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
#if RS_BWD_SMEM_REDUCE
#pragma unroll
                for (int q = 0; q < OWN; ++q)
                    if (q < own_real)
                        atomicAdd(bwd_acc + tt * NVP + own0 + q, v[q]);
                if (lane == 0)
                    atomicOr(&touched[tt >> 5], 1u << (tt & 31));
#else
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
#endif
            }
        }
    }
    if (num_batches > first_batch) {
        __syncthreads();
        flush();
    }
}

template <int CDIM>
static int launch_raster_bwd(const rs_raster_bwd_args &b, int ch_off, int ch_cnt, bool first, cudaStream_t s) {
    const int64_t grid = (int64_t)b.f.I * b.f.tile_width * b.f.tile_height;
    const bool abs_grad = b.v_means2d_abs != nullptr;
    const int nv = CDIM + 6 + (abs_grad ? 2 : 0);
    const size_t acc_bytes = RS_BWD_SMEM_REDUCE ? (size_t)RAST_THREADS * (nv | 1) * sizeof(float) : 0;
    static RsPerDevice attr[2];
    if (acc_bytes + sizeof(RastBwdSmem<CDIM>) > 47 * 1024 && !rs_dev_done(attr[abs_grad])) {
        if (abs_grad)
            RS_CUDA(cudaFuncSetAttribute(rs_raster_bwd_kernel<CDIM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_bytes));
        else
            RS_CUDA(cudaFuncSetAttribute(rs_raster_bwd_kernel<CDIM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_bytes));
        rs_dev_mark(attr[abs_grad]);
    }
    if (abs_grad)
        rs_raster_bwd_kernel<CDIM, true><<<(unsigned)grid, RAST_THREADS, acc_bytes, s>>>(b, ch_off, ch_cnt, first);
    else
        rs_raster_bwd_kernel<CDIM, false><<<(unsigned)grid, RAST_THREADS, acc_bytes, s>>>(b, ch_off, ch_cnt, first);
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
