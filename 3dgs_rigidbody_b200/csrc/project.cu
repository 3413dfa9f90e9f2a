// rs_project_fwd: rigid transform + fused EWA projection (+ optional tile counting).
// Replaces main.py:183-228 + csrc/ProjectionEWA3DGSFused.cu:15-212 (+ csrc/IntersectTile.cu:55-84 pass 1).
//
// HBM-bound streaming kernel.  One CTA (256 threads) owns RS_ISECT_BLOCK = 1024 consecutive (image, gaussian)
// elements, 4 per thread in a coalesced stride-256 pattern, so the per-CTA set-up (pose table, camera) is amortised
// and every thread has 4 independent load chains in flight.  Algorithmic bytes per element: 48 read (mean 12,
// quat 16, scale 12, opacity 4, cluster id 4) + 32 written (radii 8, mean2d 8, depth 4, conic 12) + 4 (tile count).
#include <string.h>

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


// Output of one (image, Gaussian) pair at `row`: the dense index for rs_project_fwd, the packed row for
// rs_project_packed_fwd.
__device__ __forceinline__ void rs_store_projected(const rs_project_fwd_args &a, size_t row, const RsProjected &o, bool ok,
                                                   float opac) {
    reinterpret_cast<int2 *>(a.radii)[row] = make_int2(o.rx, o.ry);
    reinterpret_cast<float2 *>(a.means2d)[row] = make_float2(o.mx, o.my);
    a.depths[row] = o.depth;
    a.conics[row * 3 + 0] = o.ca;
    a.conics[row * 3 + 1] = o.cb;
    a.conics[row * 3 + 2] = o.cc;
    if (a.compensations != nullptr)
        a.compensations[row] = o.comp;
    if (a.records != nullptr && ok) { // compositing record (see raster_fwd.cu); culled rows are never referenced
        float4 *rec = reinterpret_cast<float4 *>(a.records) + row * 2;
        rec[0] = make_float4(o.mx, o.my, opac, o.ca);
        rec[1] = make_float4(o.cb, o.cc, rs_cull_limit(o.ca, o.cb, o.cc, opac), 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Packed (COO) placement.  Chunks of 256 (image, Gaussian) pairs are taken in ticket order (so every predecessor of a
// running chunk has started) and the first packed row of a chunk is the number of visible pairs in all earlier chunks,
// found by a decoupled look-back over one 64-bit status word per chunk: {flag : 32 | count : 32}, flag 0 = not
// published, 1 = the chunk's own count, 2 = inclusive prefix.  One warp inspects 32 predecessors per round trip.
// ---------------------------------------------------------------------------------------------------------------------
#define RS_PACK_AGG (1ull << 32)
#define RS_PACK_PREFIX (2ull << 32)
struct PackCtl {
    unsigned int ticket;
    unsigned int _pad[3];
    unsigned long long state[1]; // [n_chunks]
};
struct PackSmem {
    int warp_excl[8];
    unsigned int base;
    unsigned int total;
    unsigned int chunk;
};

__device__ __forceinline__ unsigned int rs_pack_take_chunk(PackCtl *ctl, PackSmem &ps) {
    if (threadIdx.x == 0)
        ps.chunk = atomicAdd(&ctl->ticket, 1u);
    __syncthreads();
    return ps.chunk;
}

// Packed row of the calling thread (meaningful when `valid`); every thread of the 256-thread CTA must call.
// `chunk_total` / `chunk_base` are returned to all threads.
__device__ __forceinline__ unsigned int rs_pack_place(bool valid, unsigned int chunk, PackCtl *ctl, PackSmem &ps,
                                                      unsigned int &chunk_base, unsigned int &chunk_total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0)
        ps.warp_excl[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        const int v = lane < 8 ? ps.warp_excl[lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o)
                incl += n;
        }
        const unsigned int total = (unsigned int)__shfl_sync(0xffffffffu, incl, 7);
        unsigned int prefix = 0;
        volatile unsigned long long *state = ctl->state;
        if (chunk > 0) {
            if (lane == 0)
                state[chunk] = RS_PACK_AGG | total;
            int look = (int)chunk - 1;
            while (true) {
                const int j = look - (int)lane;
                unsigned long long s;
                do {
                    s = j >= 0 ? state[j] : RS_PACK_PREFIX;
                } while (__any_sync(0xffffffffu, (s >> 32) == 0ull));
                const unsigned pm = __ballot_sync(0xffffffffu, (s >> 32) == 2ull);
                unsigned int val = (unsigned int)s;
                if (pm != 0u && lane > (unsigned)(__ffs(pm) - 1))
                    val = 0u; // behind the nearest published prefix
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (pm != 0u)
                    break;
                look -= 32;
            }
        }
        if (lane == 0) {
            state[chunk] = RS_PACK_PREFIX | (unsigned long long)(prefix + total);
            ps.base = prefix;
        }
        if (lane < 8)
            ps.warp_excl[lane] = incl - v;
        if (lane == 8)
            ps.total = total;
    }
    __syncthreads();
    chunk_base = ps.base;
    chunk_total = ps.total;
    return chunk_base + (unsigned int)ps.warp_excl[warp] + (unsigned int)__popc(bal & rs_lanemask_lt());
}

struct PackOut {
    long long capacity;
    int32_t *indptr;
    long long *batch_ids, *camera_ids, *gaussian_ids;
    long long *nnz;
    PackCtl *ctl;
    unsigned int n_chunks, chunks_per_image;
};
#define RS_PACK_CHUNK 256


// The thread holding the first Gaussian of an image records where the image's rows start (its own exclusive rank, visible
// or not); the last chunk closes indptr and publishes the row count.
__device__ __forceinline__ void rs_pack_finish(const PackOut &po, unsigned int chunk, unsigned int img, unsigned int gid,
                                               bool in_range, unsigned int n_images, unsigned int row,
                                               unsigned int chunk_base, unsigned int chunk_total) {
    if (in_range && gid == 0u && po.indptr != nullptr)
        po.indptr[img] = (int32_t)row;
    if (threadIdx.x == 0 && chunk == po.n_chunks - 1) {
        if (po.indptr != nullptr)
            po.indptr[n_images] = (int32_t)(chunk_base + chunk_total);
        *po.nnz = (long long)chunk_base + chunk_total;
    }
}

// depth statistics of the rows with at least one tile (first step of the depth ordering, see depth_order.cu): warp
// reduction, CTA reduction through shared memory, then ONE atomic triple per CTA.  (One triple per warp was measured: the
// 94 k same-address RED operations of a 1 M-Gaussian launch serialise in L2 and more than double the kernel, 37 -> 82 us.)
// Every thread of the CTA must call; `red` = 3 * (blockDim.x / 32) words of shared memory.
__device__ __forceinline__ void rs_project_depth_stats(uint32_t *stats, bool counted, float depth, unsigned int *red) {
    const unsigned int k = __float_as_uint(depth);
    unsigned int inv_min = counted ? ~k : 0u, mx = counted ? k : 0u;
    unsigned int cnt = __popc(__ballot_sync(0xffffffffu, counted));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inv_min = max(inv_min, __shfl_xor_sync(0xffffffffu, inv_min, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[warp] = inv_min;
        red[n_warps + warp] = mx;
        red[2 * n_warps + warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < n_warps; ++w) {
            inv_min = max(inv_min, red[w]);
            mx = max(mx, red[n_warps + w]);
            cnt += red[2 * n_warps + w];
        }
        if (cnt != 0u) {
            atomicMax(stats + 0, inv_min);
            atomicMax(stats + 1, mx);
            atomicAdd(stats + 2, cnt);
        }
    }
}

struct ProjSmem {
    float cam[2][16];
    int sums[8];
};

template <bool HAS_RIGID, bool PACKED>
__global__ void __launch_bounds__(RS_ISECT_THREADS, 3)
rs_project_fwd_kernel(const rs_project_fwd_args a, const PackOut po) {
    extern __shared__ __align__(16) float smem_dyn[]; // pose table (HAS_RIGID only)
    __shared__ ProjSmem sm;
    __shared__ PackSmem ps; // PACKED only

    const uint32_t N = a.N, C = a.C;
    const uint64_t total = (uint64_t)a.B * C * N;
    // PACKED: one 256-pair chunk per CTA, taken in ticket order; else 1024 pairs per CTA
    const unsigned int chunk = PACKED ? rs_pack_take_chunk(po.ctl, ps) : 0u;
    const uint64_t block_base = PACKED ? (uint64_t)chunk * RS_PACK_CHUNK : (uint64_t)blockIdx.x * RS_ISECT_BLOCK;

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
    for (int it = 0; it < (PACKED ? 1 : RS_ISECT_BLOCK / RS_ISECT_THREADS); ++it) {
        const uint64_t idx_raw = block_base + (uint64_t)it * RS_ISECT_THREADS + threadIdx.x;
        const bool in_range = idx_raw < total;
        if (!PACKED && !in_range)
            break;
        // PACKED: the placement below is a block-wide step, so the out-of-range threads of the last chunk stay in the
        // loop body; they re-evaluate the last pair and drop the result
        const uint64_t idx = PACKED ? min(idx_raw, total - 1) : idx_raw;
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
        bool ok = rs_project_gaussian(mean, covar, cam, a.camera_model, (uint32_t)a.image_width,
                                      (uint32_t)a.image_height, a.eps2d, a.near_plane, a.far_plane, a.radius_clip,
                                      a.opacities != nullptr ? &opac : nullptr, a.compensations != nullptr, o);
        if (PACKED)
            ok = ok && in_range;
        if (!ok) {
            o.mx = o.my = o.depth = o.ca = o.cb = o.cc = 0.f;
            o.comp = 0.f;
        }
        size_t row = (size_t)idx;
        if (PACKED) {
            unsigned int cb, ct;
            row = rs_pack_place(ok, chunk, po.ctl, ps, cb, ct);
            rs_pack_finish(po, chunk, img, gid, in_range, (uint32_t)a.B * C, (unsigned int)row, cb, ct);
            if (!ok || (long long)row >= po.capacity)
                break; // nothing else is block-wide in PACKED mode (block_sums is rejected by the host side)
            po.batch_ids[row] = bid;
            po.camera_ids[row] = img - bid * C;
            po.gaussian_ids[row] = gid;
        }
        rs_store_projected(a, row, o, ok, opac);
        if (a.sh_coeffs != nullptr && ok)
            rs_project_sh_color(a, cam, mean, gsrc, row);
        if (a.tiles_per_gauss != nullptr) {
            int cnt;
            if (a.tile_footprints != nullptr) {
                uint4 fp;
                cnt = rs_tile_footprint(o.mx, o.my, o.rx, o.ry, o.ca, o.cb, o.cc, opac, (uint32_t)a.tile_size,
                                        (uint32_t)a.tile_width, (uint32_t)a.tile_height, fp);
                reinterpret_cast<uint4 *>(a.tile_footprints)[row] = fp;
            } else {
                cnt = rs_tile_count(o.rx, o.ry, o.mx, o.my, (uint32_t)a.tile_size, (uint32_t)a.tile_width,
                                    (uint32_t)a.tile_height);
            }
            a.tiles_per_gauss[row] = cnt;
            my_tiles += cnt;
            if (a.depth_stats != nullptr && cnt > 0) { // general path (not the frame path): per-row atomics
                const unsigned int k = __float_as_uint(o.depth);
                atomicMax(a.depth_stats + 0, ~k);
                atomicMax(a.depth_stats + 1, k);
                atomicAdd(a.depth_stats + 2, 1u);
            }
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

template <bool HAS_RIGID, bool PACKED>
__global__ void __launch_bounds__(PROJ_CHUNK, 4)
rs_project_fwd_staged_kernel(const rs_project_fwd_args a, const PackOut po) {
    extern __shared__ __align__(16) float smem_dyn[]; // pose table (HAS_RIGID only)
    __shared__ __align__(16) ProjStage st;
    __shared__ PackSmem ps; // PACKED only

    const uint32_t N = a.N, C = a.C;
    // PACKED: chunks in ticket order over a 1-D grid, image-major; else blockIdx = (chunk of the image, image)
    const unsigned int chunk = PACKED ? rs_pack_take_chunk(po.ctl, ps) : 0u;
    const uint32_t img = PACKED ? chunk / po.chunks_per_image : blockIdx.y; // bid * C + cid
    const uint32_t bid = img / C;
    const uint32_t g0 = (PACKED ? chunk - img * po.chunks_per_image : blockIdx.x) * PROJ_CHUNK;
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

    const uint32_t t = threadIdx.x;
    const bool in_range = t < n_valid;
    const uint32_t gid = g0 + t;
    RsProjected o;
    RsCam cam;
    float mean[3] = {0.f, 0.f, 0.f};
    float opac = 0.f;
    bool ok = false;
    if (in_range) {
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
        mean[0] = st.means[t * 3 + 0];
        mean[1] = st.means[t * 3 + 1];
        mean[2] = st.means[t * 3 + 2];
        const float4 q4 = reinterpret_cast<const float4 *>(st.quats)[t];
        float quat[4] = {q4.x, q4.y, q4.z, q4.w};
        float scale[3] = {st.scales[t * 3 + 0], st.scales[t * 3 + 1], st.scales[t * 3 + 2]};
        opac = st.opacities[t];
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
        ok = rs_project_gaussian(mean, covar, cam, a.camera_model, (uint32_t)a.image_width, (uint32_t)a.image_height,
                                 a.eps2d, a.near_plane, a.far_plane, a.radius_clip, &opac, a.compensations != nullptr, o);
        if (!ok) {
            o.mx = o.my = o.depth = o.ca = o.cb = o.cc = 0.f;
            o.comp = 0.f;
        }
    }
    size_t row = (size_t)img * N + gid;
    if (PACKED) {
        unsigned int cb, ct;
        row = rs_pack_place(ok, chunk, po.ctl, ps, cb, ct);
        rs_pack_finish(po, chunk, img, gid, in_range, (uint32_t)a.B * C, (unsigned int)row, cb, ct);
        if (!ok || (long long)row >= po.capacity)
            return;
        po.batch_ids[row] = bid;
        po.camera_ids[row] = img - bid * C;
        po.gaussian_ids[row] = gid;
    }
    int cnt = 0;
    if constexpr (!PACKED) { // (every thread of the CTA reaches this point when not PACKED)
        if (a.tile_footprints != nullptr) { // tight tile lists, computed warp-wide
            __shared__ RsFootWarp foot[PROJ_CHUNK / 32];
            uint4 fp;
            cnt = rs_tile_footprint_warp(in_range && ok, o.mx, o.my, o.rx, o.ry, o.ca, o.cb, o.cc, opac, (uint32_t)a.tile_size,
                                         (uint32_t)a.tile_width, (uint32_t)a.tile_height, fp, foot[threadIdx.x >> 5]);
            if (in_range) {
                reinterpret_cast<uint4 *>(a.tile_footprints)[row] = fp;
                a.tiles_per_gauss[row] = cnt;
            }
        }
    }
    if (in_range) {
        rs_store_projected(a, row, o, ok, opac);
        if (a.sh_coeffs != nullptr && ok)
            rs_project_sh_color(a, cam, mean, src0 + t, row);
        if (a.tiles_per_gauss != nullptr && (PACKED || a.tile_footprints == nullptr)) {
            if (a.tile_footprints != nullptr) {
                uint4 fp;
                cnt = rs_tile_footprint(o.mx, o.my, o.rx, o.ry, o.ca, o.cb, o.cc, opac, (uint32_t)a.tile_size,
                                        (uint32_t)a.tile_width, (uint32_t)a.tile_height, fp);
                reinterpret_cast<uint4 *>(a.tile_footprints)[row] = fp;
            } else {
                cnt = rs_tile_count(o.rx, o.ry, o.mx, o.my, (uint32_t)a.tile_size, (uint32_t)a.tile_width,
                                    (uint32_t)a.tile_height);
            }
            a.tiles_per_gauss[row] = cnt;
        }
    }
    if (!PACKED && a.depth_stats != nullptr) { // (every thread of the CTA reaches this point when not PACKED)
        __shared__ unsigned int stats_red[3 * (PROJ_CHUNK / 32)];
        rs_project_depth_stats(a.depth_stats, cnt > 0, o.depth, stats_red);
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// validation + launch shared by the dense and the packed entry points (po == nullptr: dense)
static int rs_project_launch(const rs_project_fwd_args *a, PackOut *po, cudaStream_t s, const char *who) {
    RS_CHECK(a->B >= 0 && a->C >= 0 && a->N >= 0, "%s: negative sizes", who);
    RS_CHECK(a->camera_model == RS_PINHOLE || a->camera_model == RS_ORTHO || a->camera_model == RS_FISHEYE,
             "%s: unsupported camera model %d (ftheta is only available through the reference's UT path)", who,
             a->camera_model);
    RS_CHECK((a->covars != nullptr) != (a->quats != nullptr && a->scales != nullptr),
             "%s: exactly one of covars or (quats, scales) must be given", who);
    const int64_t total = (int64_t)a->B * a->C * a->N;
    if (total == 0)
        return 0;
    RS_CHECK(total < (int64_t)1 << 31, "%s: B*C*N = %lld exceeds int32 indexing", who, (long long)total);
    RS_CHECK(a->means && a->viewmats && a->Ks && a->radii && a->means2d && a->depths && a->conics,
             "%s: null required pointer", who);
    if (a->tiles_per_gauss != nullptr)
        RS_CHECK(a->tile_size > 0 && a->tile_width > 0 && a->tile_height > 0,
                 "%s: tile geometry required for fused tile counting", who);
    RS_CHECK(a->tile_footprints == nullptr || (a->tiles_per_gauss != nullptr && a->opacities != nullptr &&
                                               aligned16(a->tile_footprints)),
             "%s: tile_footprints needs tiles_per_gauss, opacities and a 16-byte aligned array", who);
    RS_CHECK(a->tile_footprints == nullptr || (a->tile_width < 65536 && a->tile_height < 65536),
             "%s: tile_footprints: more than 65535 tiles per axis", who);
    RS_CHECK(a->depth_stats == nullptr || (a->tiles_per_gauss != nullptr && po == nullptr),
             "%s: depth_stats needs tiles_per_gauss and dense (not packed) rows", who);
    if (a->sh_coeffs != nullptr)
        RS_CHECK(a->sh_colors != nullptr && a->sh_degree >= 0 && a->sh_degree <= 4 &&
                     (a->sh_degree + 1) * (a->sh_degree + 1) <= a->sh_K,
                 "%s: bad SH arguments (degree %d, K %d)", who, a->sh_degree, a->sh_K);
    const bool rigid = a->rigid.cluster_ids != nullptr;
    if (rigid)
        RS_CHECK(a->rigid.body_quats && a->rigid.body_trans && a->rigid.K > 0, "%s: rigid table incomplete (K=%d)", who,
                 a->rigid.K);
    const size_t pose_smem =
        rigid && a->rigid.K <= RS_MAX_SMEM_BODIES ? (size_t)a->rigid.K * RS_BODY_FLOATS * sizeof(float) : 0;
    const bool packed = po != nullptr;
    PackOut none;
    memset(&none, 0, sizeof(none));
    // fast path: TMA-staged inputs (see rs_project_fwd_staged_kernel).  block_sums are only needed by the unsorted
    // rs_isect_emit path, which takes them from rs_isect_count, so the staged kernel does not produce them.
    if (a->quats != nullptr && a->opacities != nullptr && a->block_sums == nullptr && aligned16(a->means) &&
        aligned16(a->quats) && aligned16(a->scales) && aligned16(a->opacities) &&
        (!rigid || aligned16(a->rigid.cluster_ids)) && (a->B == 1 || a->N % 4 == 0) && (int64_t)a->B * a->C <= 65535) {
        const unsigned chunks_per_image = (unsigned)((a->N + PROJ_CHUNK - 1) / PROJ_CHUNK);
        if (rigid) {
            static RsPerDevice attr_set; // pose table + input stage can exceed the 48 KB default
            if (!rs_dev_done(attr_set)) {
                const int bytes = RS_MAX_SMEM_BODIES * RS_BODY_FLOATS * (int)sizeof(float);
                RS_CUDA(cudaFuncSetAttribute(rs_project_fwd_staged_kernel<true, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                RS_CUDA(cudaFuncSetAttribute(rs_project_fwd_staged_kernel<true, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                rs_dev_mark(attr_set);
            }
        }
        if (packed) {
            po->chunks_per_image = chunks_per_image;
            po->n_chunks = chunks_per_image * (unsigned)(a->B * a->C);
            if (rigid)
                rs_project_fwd_staged_kernel<true, true><<<po->n_chunks, PROJ_CHUNK, pose_smem, s>>>(*a, *po);
            else
                rs_project_fwd_staged_kernel<false, true><<<po->n_chunks, PROJ_CHUNK, 0, s>>>(*a, *po);
        } else {
            const dim3 grid2(chunks_per_image, (unsigned)(a->B * a->C));
            if (rigid)
                rs_project_fwd_staged_kernel<true, false><<<grid2, PROJ_CHUNK, pose_smem, s>>>(*a, none);
            else
                rs_project_fwd_staged_kernel<false, false><<<grid2, PROJ_CHUNK, 0, s>>>(*a, none);
        }
        RS_LAUNCH_CHECK("rs_project_fwd_staged_kernel");
        return 0;
    }
    if (packed) {
        po->chunks_per_image = 0; // chunks run over the flat (image, Gaussian) index
        po->n_chunks = (unsigned)((total + RS_PACK_CHUNK - 1) / RS_PACK_CHUNK);
        if (rigid)
            rs_project_fwd_kernel<true, true><<<po->n_chunks, RS_ISECT_THREADS, pose_smem, s>>>(*a, *po);
        else
            rs_project_fwd_kernel<false, true><<<po->n_chunks, RS_ISECT_THREADS, 0, s>>>(*a, *po);
    } else {
        const int grid = rs_isect_num_blocks(total);
        if (rigid)
            rs_project_fwd_kernel<true, false><<<grid, RS_ISECT_THREADS, pose_smem, s>>>(*a, none);
        else
            rs_project_fwd_kernel<false, false><<<grid, RS_ISECT_THREADS, 0, s>>>(*a, none);
    }
    RS_LAUNCH_CHECK("rs_project_fwd_kernel");
    return 0;
}

extern "C" int rs_project_fwd(const rs_project_fwd_args *a, rs_stream_t stream) {
    RS_CHECK(a != nullptr, "rs_project_fwd: null args");
    return rs_project_launch(a, nullptr, (cudaStream_t)stream, "rs_project_fwd");
}

// ticket + one status word per 256-pair chunk (the staged kernel chunks per image, the general one over the flat index;
// the former never has fewer chunks)
extern "C" uint64_t rs_project_packed_workspace_bytes(int32_t B, int32_t C, int32_t N) {
    const uint64_t chunks = (uint64_t)((N + PROJ_CHUNK - 1) / PROJ_CHUNK) * (uint64_t)B * (uint64_t)C;
    return sizeof(PackCtl) + (chunks + 1) * sizeof(unsigned long long);
}

extern "C" int rs_project_packed_fwd(const rs_project_packed_fwd_args *pa, rs_stream_t stream) {
    RS_CHECK(pa != nullptr, "rs_project_packed_fwd: null args");
    const rs_project_fwd_args *a = &pa->proj;
    RS_CHECK(pa->nnz != nullptr, "rs_project_packed_fwd: nnz is required");
    RS_CHECK(a->block_sums == nullptr, "rs_project_packed_fwd: block_sums is not available for packed rows");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = (int64_t)a->B * a->C * a->N;
    if (total <= 0) { // no rows: indptr all zero
        RS_CHECK(a->B >= 0 && a->C >= 0 && a->N >= 0, "rs_project_packed_fwd: negative sizes");
        RS_CUDA(cudaMemsetAsync(pa->nnz, 0, sizeof(int64_t), s));
        if (pa->indptr != nullptr)
            RS_CUDA(cudaMemsetAsync(pa->indptr, 0, ((size_t)a->B * a->C + 1) * sizeof(int32_t), s));
        return 0;
    }
    RS_CHECK(pa->capacity >= 0 && pa->workspace != nullptr, "rs_project_packed_fwd: workspace / capacity missing");
    RS_CHECK(pa->capacity == 0 || (pa->batch_ids && pa->camera_ids && pa->gaussian_ids),
             "rs_project_packed_fwd: id outputs are required");
    RS_CUDA(cudaMemsetAsync(pa->workspace, 0, rs_project_packed_workspace_bytes(a->B, a->C, a->N), s));
    PackOut po;
    po.capacity = pa->capacity;
    po.indptr = pa->indptr;
    po.batch_ids = (long long *)pa->batch_ids;
    po.camera_ids = (long long *)pa->camera_ids;
    po.gaussian_ids = (long long *)pa->gaussian_ids;
    po.nnz = (long long *)pa->nnz;
    po.ctl = (PackCtl *)pa->workspace;
    po.n_chunks = po.chunks_per_image = 0;
    return rs_project_launch(a, &po, s, "rs_project_packed_fwd");
}
