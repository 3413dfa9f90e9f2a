// rs_project_fwd: rigid transform + fused EWA projection (+ optional tile counting).
// Replaces main.py:183-228 + csrc/ProjectionEWA3DGSFused.cu:15-212 (+ csrc/IntersectTile.cu:55-84 pass 1).
//
// HBM-bound streaming kernel.  One CTA (256 threads) owns RS_ISECT_BLOCK = 1024 consecutive (image, gaussian)
// elements, 4 per thread in a coalesced stride-256 pattern, so the per-CTA set-up (pose table, camera) is amortised
// and every thread has 4 independent load chains in flight.  Algorithmic bytes per element: 48 read (mean 12,
// quat 16, scale 12, opacity 4, cluster id 4) + 32 written (radii 8, mean2d 8, depth 4, conic 12) + 4 (tile count).
#include "project_math.cuh"
#include "sh_math.cuh"

// colour of one visible Gaussian from its SH coefficients, view direction = moved mean - camera origin (-R^T t)
__device__ __forceinline__ void rs_project_sh_color(const rs_project_fwd_args &a, const RsCam &cam, const float mean[3],
                                                    size_t gsrc, size_t idx) {
    const float ox = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    const float oy = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    const float oz = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    const float dx = mean[0] - ox, dy = mean[1] - oy, dz = mean[2] - oz;
    const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
    float B[25], c[3];
    rs_sh_basis(a.sh_degree, dx * inorm, dy * inorm, dz * inorm, B);
    rs_sh_dot(B, (a.sh_degree + 1) * (a.sh_degree + 1), a.sh_coeffs + gsrc * (size_t)a.sh_K * 3, c);
    a.sh_colors[idx * 3 + 0] = fmaxf(c[0] + 0.5f, 0.f);
    a.sh_colors[idx * 3 + 1] = fmaxf(c[1] + 0.5f, 0.f);
    a.sh_colors[idx * 3 + 2] = fmaxf(c[2] + 0.5f, 0.f);
}

struct ProjSmem {
    float cam[2][16];
    int sums[8];
};

template <bool HAS_RIGID>
__global__ void __launch_bounds__(RS_ISECT_THREADS, 3)
rs_project_fwd_kernel(const rs_project_fwd_args a) {
    extern __shared__ __align__(16) float smem_dyn[]; // pose table (HAS_RIGID only)
    __shared__ ProjSmem sm;

    const uint32_t N = a.N, C = a.C;
    const uint64_t total = (uint64_t)a.B * C * N;
    const uint64_t block_base = (uint64_t)blockIdx.x * RS_ISECT_BLOCK;

    if (HAS_RIGID)
        rs_load_pose_table(a.rigid, smem_dyn);
    // cache the (at most two) cameras most elements of this CTA use
    const uint32_t img0 = (uint32_t)(block_base / N);
    if (threadIdx.x < 32) {
        int which = threadIdx.x >> 4, i = threadIdx.x & 15;
        uint32_t img = img0 + which;
        if (img < (uint32_t)a.B * C) {
            // 12 viewmat entries + fx, fy, cx, cy
            float v;
            if (i < 12)
                v = a.viewmats[(size_t)img * 16 + i];
            else {
                const int kidx[4] = {0, 4, 2, 5};
                v = a.Ks[(size_t)img * 9 + kidx[i - 12]];
            }
            sm.cam[which][i] = v;
        }
    }
    __syncthreads();

    int my_tiles = 0;
#pragma unroll 1
    for (int it = 0; it < RS_ISECT_BLOCK / RS_ISECT_THREADS; ++it) {
        const uint64_t idx = block_base + (uint64_t)it * RS_ISECT_THREADS + threadIdx.x;
        if (idx >= total)
            break;
        // (image, gaussian) of this element; the common single-image case needs no division at all and totals below
        // 2^32 avoid the 64-bit divide
        uint32_t img, gid;
        if (total <= N) {
            img = 0;
            gid = (uint32_t)idx;
        } else if (total <= 0xffffffffull) {
            img = (uint32_t)idx / N;
            gid = (uint32_t)idx - img * N;
        } else {
            img = (uint32_t)(idx / N); // bid * C + cid
            gid = (uint32_t)(idx - (uint64_t)img * N);
        }
        const uint32_t bid = (C == 1) ? img : img / C;
        const size_t gsrc = (size_t)bid * N + gid; // row in the per-batch Gaussian arrays

        RsCam cam;
        if (img - img0 < 2u) {
            const float *cs = sm.cam[img - img0];
            cam.R[0] = cs[0];
            cam.R[1] = cs[1];
            cam.R[2] = cs[2];
            cam.t[0] = cs[3];
            cam.R[3] = cs[4];
            cam.R[4] = cs[5];
            cam.R[5] = cs[6];
            cam.t[1] = cs[7];
            cam.R[6] = cs[8];
            cam.R[7] = cs[9];
            cam.R[8] = cs[10];
            cam.t[2] = cs[11];
            cam.fx = cs[12];
            cam.fy = cs[13];
            cam.cx = cs[14];
            cam.cy = cs[15];
        } else {
            rs_load_cam(a.viewmats + (size_t)img * 16, a.Ks + (size_t)img * 9, cam);
        }

        float mean[3];
        mean[0] = a.means[gsrc * 3 + 0];
        mean[1] = a.means[gsrc * 3 + 1];
        mean[2] = a.means[gsrc * 3 + 2];
        float quat[4] = {1.f, 0.f, 0.f, 0.f};
        float covar[9];
        const bool has_quat = a.covars == nullptr;
        if (has_quat) {
            const float4 q4 = *reinterpret_cast<const float4 *>(a.quats + gsrc * 4);
            quat[0] = q4.x;
            quat[1] = q4.y;
            quat[2] = q4.z;
            quat[3] = q4.w;
        }
        float body[RS_BODY_FLOATS];
        int k = -1;
        if (HAS_RIGID)
            k = rs_rigid_transform(a.rigid, smem_dyn, gid, mean, quat, has_quat, has_quat ? nullptr : body);
        if (has_quat) {
            float scale[3];
            scale[0] = a.scales[gsrc * 3 + 0];
            scale[1] = a.scales[gsrc * 3 + 1];
            scale[2] = a.scales[gsrc * 3 + 2];
            rs_quat_scale_to_covar(quat, scale, covar, nullptr);
        } else {
            const float *cv = a.covars + gsrc * 6;
            covar[0] = cv[0];
            covar[1] = cv[1];
            covar[2] = cv[2];
            covar[3] = cv[1];
            covar[4] = cv[3];
            covar[5] = cv[4];
            covar[6] = cv[2];
            covar[7] = cv[4];
            covar[8] = cv[5];
            if (HAS_RIGID && k >= 0) { // Sigma' = R_k Sigma R_k^T
                float tmp[9];
                rs_mm3(body, covar, tmp);
                rs_mm3_nt(tmp, body, covar);
            }
        }
        float opac = 0.f;
        if (a.opacities != nullptr)
            opac = a.opacities[gsrc];

        RsProjected o;
        const bool ok = rs_project_gaussian(mean, covar, cam, a.camera_model, (uint32_t)a.image_width,
                                            (uint32_t)a.image_height, a.eps2d, a.near_plane, a.far_plane,
                                            a.radius_clip, a.opacities != nullptr ? &opac : nullptr,
                                            a.compensations != nullptr, o);
        if (!ok) {
            o.mx = o.my = o.depth = o.ca = o.cb = o.cc = 0.f;
            o.comp = 0.f;
        }
        reinterpret_cast<int2 *>(a.radii)[idx] = make_int2(o.rx, o.ry);
        reinterpret_cast<float2 *>(a.means2d)[idx] = make_float2(o.mx, o.my);
        a.depths[idx] = o.depth;
        a.conics[idx * 3 + 0] = o.ca;
        a.conics[idx * 3 + 1] = o.cb;
        a.conics[idx * 3 + 2] = o.cc;
        if (a.compensations != nullptr)
            a.compensations[idx] = o.comp;
        if (a.records != nullptr && ok) { // compositing record (see raster_fwd.cu); culled rows are never referenced
            float4 *rec = reinterpret_cast<float4 *>(a.records) + idx * 2;
            rec[0] = make_float4(o.mx, o.my, opac, o.ca);
            rec[1] = make_float4(o.cb, o.cc, rs_cull_limit(o.ca, o.cb, o.cc, opac), 0.f);
        }
        if (a.sh_coeffs != nullptr && ok)
            rs_project_sh_color(a, cam, mean, gsrc, (size_t)idx);
        if (a.tiles_per_gauss != nullptr) {
            int cnt = rs_tile_count(o.rx, o.ry, o.mx, o.my, (uint32_t)a.tile_size, (uint32_t)a.tile_width,
                                    (uint32_t)a.tile_height);
            a.tiles_per_gauss[idx] = cnt;
            my_tiles += cnt;
        }
    }
    if (a.block_sums != nullptr) {
        int s = rs_block_sum_256(my_tiles, sm.sums);
        if (threadIdx.x == 0)
            a.block_sums[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path (quats + scales + opacities, 16-byte aligned arrays): one CTA = 256 consecutive Gaussians of ONE image.
// The CTA's slices of the AoS inputs -- means (3072 B), quats (4096 B), scales (3072 B), opacities (1024 B), cluster ids
// (1024 B) -- are contiguous in HBM, so one elected thread fetches them with five TMA bulk copies (cp.async.bulk,
// SASS UBLKCP) that complete on an mbarrier; every thread then reads its own Gaussian from shared memory (stride-3 word
// access is bank-conflict free).  No thread issues a 12-byte strided global load, nothing is staged through registers,
// and the whole input of the CTA is in flight after one instruction per array.
// ---------------------------------------------------------------------------------------------------------------------
#define PROJ_CHUNK 256
struct ProjStage {
    float means[PROJ_CHUNK * 3];
    float quats[PROJ_CHUNK * 4];
    float scales[PROJ_CHUNK * 3];
    float opacities[PROJ_CHUNK];
    int32_t ids[PROJ_CHUNK];
    float cam[16];
    int sums[8];
    unsigned long long bar;
};

__device__ __forceinline__ void rs_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <bool HAS_RIGID>
__global__ void __launch_bounds__(PROJ_CHUNK, 3)
rs_project_fwd_staged_kernel(const rs_project_fwd_args a) {
    extern __shared__ __align__(16) float smem_dyn[]; // pose table (HAS_RIGID only)
    __shared__ __align__(16) ProjStage st;

    const uint32_t N = a.N, C = a.C;
    const uint32_t img = blockIdx.y; // bid * C + cid
    const uint32_t bid = img / C;
    const uint32_t g0 = blockIdx.x * PROJ_CHUNK;
    const uint32_t n_valid = min((uint32_t)PROJ_CHUNK, N - g0);
    const size_t src0 = (size_t)bid * N + g0; // first row of this chunk in the per-batch Gaussian arrays
    const bool full = n_valid == PROJ_CHUNK;
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&st.bar);

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;\n" ::"r"(bar_addr) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (full) {
        if (threadIdx.x == 0) {
            uint32_t tx = PROJ_CHUNK * (12 + 16 + 12 + 4) + (HAS_RIGID ? PROJ_CHUNK * 4 : 0);
            asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared.b64 t, [%0], %1;\n\t}\n" ::"r"(bar_addr), "r"(tx)
                         : "memory");
            rs_bulk_g2s(st.means, a.means + src0 * 3, PROJ_CHUNK * 12, &st.bar);
            rs_bulk_g2s(st.quats, a.quats + src0 * 4, PROJ_CHUNK * 16, &st.bar);
            rs_bulk_g2s(st.scales, a.scales + src0 * 3, PROJ_CHUNK * 12, &st.bar);
            rs_bulk_g2s(st.opacities, a.opacities + src0, PROJ_CHUNK * 4, &st.bar);
            if (HAS_RIGID)
                rs_bulk_g2s(st.ids, a.rigid.cluster_ids + g0, PROJ_CHUNK * 4, &st.bar);
        }
    } else { // ragged last chunk: sizes are not multiples of 16 bytes, stage it with plain loads
        for (uint32_t i = threadIdx.x; i < n_valid * 3; i += PROJ_CHUNK) {
            st.means[i] = a.means[src0 * 3 + i];
            st.scales[i] = a.scales[src0 * 3 + i];
        }
        for (uint32_t i = threadIdx.x; i < n_valid * 4; i += PROJ_CHUNK)
            st.quats[i] = a.quats[src0 * 4 + i];
        for (uint32_t i = threadIdx.x; i < n_valid; i += PROJ_CHUNK) {
            st.opacities[i] = a.opacities[src0 + i];
            st.ids[i] = HAS_RIGID ? a.rigid.cluster_ids[g0 + i] : -1;
        }
    }
    // overlapped with the bulk copies: pose table and camera
    if (HAS_RIGID)
        rs_load_pose_table(a.rigid, smem_dyn);
    if (threadIdx.x < 16) {
        const int i = threadIdx.x;
        const int kidx[4] = {0, 4, 2, 5};
        st.cam[i] = (i < 12) ? a.viewmats[(size_t)img * 16 + i] : a.Ks[(size_t)img * 9 + kidx[i - 12]];
    }
    __syncthreads();
    if (full) {
        unsigned ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok)
                         : "r"(bar_addr)
                         : "memory");
    }

    int my_tiles = 0;
    const uint32_t t = threadIdx.x;
    if (t < n_valid) {
        const uint32_t gid = g0 + t;
        const size_t idx = (size_t)img * N + gid;
        RsCam cam;
        cam.R[0] = st.cam[0];
        cam.R[1] = st.cam[1];
        cam.R[2] = st.cam[2];
        cam.t[0] = st.cam[3];
        cam.R[3] = st.cam[4];
        cam.R[4] = st.cam[5];
        cam.R[5] = st.cam[6];
        cam.t[1] = st.cam[7];
        cam.R[6] = st.cam[8];
        cam.R[7] = st.cam[9];
        cam.R[8] = st.cam[10];
        cam.t[2] = st.cam[11];
        cam.fx = st.cam[12];
        cam.fy = st.cam[13];
        cam.cx = st.cam[14];
        cam.cy = st.cam[15];
        float mean[3] = {st.means[t * 3 + 0], st.means[t * 3 + 1], st.means[t * 3 + 2]};
        const float4 q4 = reinterpret_cast<const float4 *>(st.quats)[t];
        float quat[4] = {q4.x, q4.y, q4.z, q4.w};
        float scale[3] = {st.scales[t * 3 + 0], st.scales[t * 3 + 1], st.scales[t * 3 + 2]};
        const float opac = st.opacities[t];
        if (HAS_RIGID) {
            const int k = st.ids[t];
            if (k >= 0 && k < a.rigid.K) {
                if (a.rigid.K <= RS_MAX_SMEM_BODIES) {
                    rs_apply_body(smem_dyn + RS_BODY_FLOATS * k, mean, quat, true);
                } else {
                    float local[RS_BODY_FLOATS];
                    rs_make_body(a.rigid, k, local);
                    rs_apply_body(local, mean, quat, true);
                }
            }
        }
        float covar[9];
        rs_quat_scale_to_covar(quat, scale, covar, nullptr);
        RsProjected o;
        const bool ok = rs_project_gaussian(mean, covar, cam, a.camera_model, (uint32_t)a.image_width,
                                            (uint32_t)a.image_height, a.eps2d, a.near_plane, a.far_plane, a.radius_clip,
                                            &opac, a.compensations != nullptr, o);
        if (!ok) {
            o.mx = o.my = o.depth = o.ca = o.cb = o.cc = 0.f;
            o.comp = 0.f;
        }
        reinterpret_cast<int2 *>(a.radii)[idx] = make_int2(o.rx, o.ry);
        reinterpret_cast<float2 *>(a.means2d)[idx] = make_float2(o.mx, o.my);
        a.depths[idx] = o.depth;
        a.conics[idx * 3 + 0] = o.ca;
        a.conics[idx * 3 + 1] = o.cb;
        a.conics[idx * 3 + 2] = o.cc;
        if (a.compensations != nullptr)
            a.compensations[idx] = o.comp;
        if (a.records != nullptr && ok) {
            float4 *rec = reinterpret_cast<float4 *>(a.records) + idx * 2;
            rec[0] = make_float4(o.mx, o.my, opac, o.ca);
            rec[1] = make_float4(o.cb, o.cc, rs_cull_limit(o.ca, o.cb, o.cc, opac), 0.f);
        }
        if (a.sh_coeffs != nullptr && ok)
            rs_project_sh_color(a, cam, mean, src0 + t, idx);
        if (a.tiles_per_gauss != nullptr) {
            my_tiles = rs_tile_count(o.rx, o.ry, o.mx, o.my, (uint32_t)a.tile_size, (uint32_t)a.tile_width,
                                     (uint32_t)a.tile_height);
            a.tiles_per_gauss[idx] = my_tiles;
        }
    }
    (void)my_tiles;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int rs_project_fwd(const rs_project_fwd_args *a, rs_stream_t stream) {
    RS_CHECK(a != nullptr, "rs_project_fwd: null args");
    RS_CHECK(a->B >= 0 && a->C >= 0 && a->N >= 0, "rs_project_fwd: negative sizes");
    RS_CHECK(a->camera_model == RS_PINHOLE || a->camera_model == RS_ORTHO || a->camera_model == RS_FISHEYE,
             "rs_project_fwd: unsupported camera model %d (ftheta is only available through the reference's UT path)",
             a->camera_model);
    RS_CHECK((a->covars != nullptr) != (a->quats != nullptr && a->scales != nullptr),
             "rs_project_fwd: exactly one of covars or (quats, scales) must be given");
    const int64_t total = (int64_t)a->B * a->C * a->N;
    if (total == 0)
        return 0;
    RS_CHECK(total < (int64_t)1 << 31, "rs_project_fwd: B*C*N = %lld exceeds int32 indexing", (long long)total);
    RS_CHECK(a->means && a->viewmats && a->Ks && a->radii && a->means2d && a->depths && a->conics,
             "rs_project_fwd: null required pointer");
    if (a->tiles_per_gauss != nullptr)
        RS_CHECK(a->tile_size > 0 && a->tile_width > 0 && a->tile_height > 0,
                 "rs_project_fwd: tile geometry required for fused tile counting");
    if (a->sh_coeffs != nullptr)
        RS_CHECK(a->sh_colors != nullptr && a->sh_degree >= 0 && a->sh_degree <= 4 &&
                     (a->sh_degree + 1) * (a->sh_degree + 1) <= a->sh_K,
                 "rs_project_fwd: bad SH arguments (degree %d, K %d)", a->sh_degree, a->sh_K);
    const bool rigid = a->rigid.cluster_ids != nullptr;
    if (rigid)
        RS_CHECK(a->rigid.body_quats && a->rigid.body_trans && a->rigid.K > 0,
                 "rs_project_fwd: rigid table incomplete (K=%d)", a->rigid.K);
    cudaStream_t s = (cudaStream_t)stream;
    // fast path: TMA-staged inputs (see rs_project_fwd_staged_kernel).  block_sums are only needed by the unsorted
    // rs_isect_emit path, which takes them from rs_isect_count, so the staged kernel does not produce them.
    if (a->quats != nullptr && a->opacities != nullptr && a->block_sums == nullptr && aligned16(a->means) &&
        aligned16(a->quats) && aligned16(a->scales) && aligned16(a->opacities) &&
        (!rigid || aligned16(a->rigid.cluster_ids)) && (a->B == 1 || a->N % 4 == 0) && (int64_t)a->B * a->C <= 65535) {
        const dim3 grid2((unsigned)((a->N + PROJ_CHUNK - 1) / PROJ_CHUNK), (unsigned)(a->B * a->C));
        if (rigid) {
            size_t smem = a->rigid.K <= RS_MAX_SMEM_BODIES ? (size_t)a->rigid.K * RS_BODY_FLOATS * sizeof(float) : 0;
            static bool attr_set = false; // pose table + input stage can exceed the 48 KB default
            if (!attr_set) {
                RS_CUDA(cudaFuncSetAttribute(rs_project_fwd_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             RS_MAX_SMEM_BODIES * RS_BODY_FLOATS * (int)sizeof(float)));
                attr_set = true;
            }
            rs_project_fwd_staged_kernel<true><<<grid2, PROJ_CHUNK, smem, s>>>(*a);
        } else {
            rs_project_fwd_staged_kernel<false><<<grid2, PROJ_CHUNK, 0, s>>>(*a);
        }
        RS_LAUNCH_CHECK("rs_project_fwd_staged_kernel");
        return 0;
    }
    const int grid = rs_isect_num_blocks(total);
    if (rigid) {
        size_t smem = a->rigid.K <= RS_MAX_SMEM_BODIES ? (size_t)a->rigid.K * RS_BODY_FLOATS * sizeof(float) : 0;
        rs_project_fwd_kernel<true><<<grid, RS_ISECT_THREADS, smem, s>>>(*a);
    } else {
        rs_project_fwd_kernel<false><<<grid, RS_ISECT_THREADS, 0, s>>>(*a);
    }
    RS_LAUNCH_CHECK("rs_project_fwd_kernel");
    return 0;
}
