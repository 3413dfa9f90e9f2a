// rs_project_fwd: rigid transform + fused EWA projection (+ optional tile counting).
// Replaces main.py:183-228 + csrc/ProjectionEWA3DGSFused.cu:15-212 (+ csrc/IntersectTile.cu:55-84 pass 1).
//
// HBM-bound streaming kernel.  One CTA (256 threads) owns RS_ISECT_BLOCK = 1024 consecutive (image, gaussian)
// elements, 4 per thread in a coalesced stride-256 pattern, so the per-CTA set-up (pose table, camera) is amortised
// and every thread has 4 independent load chains in flight.  Algorithmic bytes per element: 48 read (mean 12,
// quat 16, scale 12, opacity 4, cluster id 4) + 32 written (radii 8, mean2d 8, depth 4, conic 12) + 4 (tile count).
#include "project_math.cuh"

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
    const bool rigid = a->rigid.cluster_ids != nullptr;
    if (rigid)
        RS_CHECK(a->rigid.body_quats && a->rigid.body_trans && a->rigid.K > 0,
                 "rs_project_fwd: rigid table incomplete (K=%d)", a->rigid.K);
    const int grid = rs_isect_num_blocks(total);
    cudaStream_t s = (cudaStream_t)stream;
    if (rigid) {
        size_t smem = a->rigid.K <= RS_MAX_SMEM_BODIES ? (size_t)a->rigid.K * RS_BODY_FLOATS * sizeof(float) : 0;
        rs_project_fwd_kernel<true><<<grid, RS_ISECT_THREADS, smem, s>>>(*a);
    } else {
        rs_project_fwd_kernel<false><<<grid, RS_ISECT_THREADS, 0, s>>>(*a);
    }
    RS_LAUNCH_CHECK("rs_project_fwd_kernel");
    return 0;
}
