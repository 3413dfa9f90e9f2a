// Tile intersection: count, scan, key emission, offsets.
// Replaces csrc/Intersect.cpp:15-168 + csrc/IntersectTile.cu:23-292 (the cub sort is in sort.cu).
//
// The reference runs the per-Gaussian kernel twice around an `at::cumsum` + `.item()` host sync and lets ONE thread
// write all tiles of a Gaussian (serial double loop, IntersectTile.cu:102-113).  Here:
//   count  : 1 thread / element, writes tiles_per_gauss and one partial sum per 1024-element block
//   scan   : one CTA turns the block sums into exclusive offsets and the device-side total (no host sync needed)
//   emit   : (operator path, unsorted 64-bit keys) a CTA re-scans its 1024 counts in shared memory and its 256 threads
//            write the block's intersections cooperatively: output slot j finds its owner by binary search, so stores
//            are fully coalesced and a Gaussian that covers thousands of tiles is spread over the whole CTA.
//   sorted : (frame path and intersect_tile(sort=True)) rs_isect_sorted further down -- depth order of the elements
//            (depth_order.cu), emission of 32-bit (image | tile) keys in that order through a shared-memory window
//            (rs_bin_emit_kernel; optionally only the tiles of the projection's tile footprints), two stable radix
//            passes on the tile bits (sort.cu), offsets.
// All of it is HBM-bound integer work: 20 B read per element + 12 B written per intersection.
#include "common.cuh"
#include "sort_ws.cuh"

extern "C" int32_t rs_isect_num_blocks(int64_t n_elems) { return (int32_t)((n_elems + RS_ISECT_BLOCK - 1) / RS_ISECT_BLOCK); }

__global__ void __launch_bounds__(RS_ISECT_THREADS) rs_isect_count_kernel(const rs_isect_args a) {
    __shared__ int sums[8];
    const int64_t base = (int64_t)blockIdx.x * RS_ISECT_BLOCK;
    int mine = 0;
#pragma unroll
    for (int it = 0; it < RS_ISECT_BLOCK / RS_ISECT_THREADS; ++it) {
        const int64_t idx = base + it * RS_ISECT_THREADS + threadIdx.x;
        if (idx < a.n_elems) {
            const int2 r = reinterpret_cast<const int2 *>(a.radii)[idx];
            int cnt = 0;
            if (r.x > 0 && r.y > 0) {
                const float2 m = reinterpret_cast<const float2 *>(a.means2d)[idx];
                cnt = rs_tile_count(r.x, r.y, m.x, m.y, (uint32_t)a.tile_size, (uint32_t)a.tile_width,
                                    (uint32_t)a.tile_height);
            }
            a.tiles_per_gauss[idx] = cnt;
            mine += cnt;
        }
    }
    int s = rs_block_sum_256(mine, sums);
    if (threadIdx.x == 0)
        a.block_sums[blockIdx.x] = s;
}

// Tile counts + tile footprints of operator-level rows in one pass (rs_isect_footprints): what the projection kernel
// writes for the frame path (rs_project_fwd_args.tile_footprints), for rows that come from somewhere else -- the receive
// arrays of the splat exchange.  With conics + opacities the masks are tight (only tiles where the splat can reach
// alpha >= 1/255), without them every tile of the bounding rectangle is listed (exactly the reference's lists).
__global__ void __launch_bounds__(RS_ISECT_THREADS)
rs_isect_footprints_kernel(const rs_isect_args a, const float *__restrict__ conics, const float *__restrict__ opacities,
                           uint4 *__restrict__ footprints) {
    __shared__ RsFootWarp foot[RS_ISECT_THREADS / 32];
    __shared__ int sums[8];
    const int64_t base = (int64_t)blockIdx.x * RS_ISECT_BLOCK;
    int mine = 0;
#pragma unroll 1
    for (int it = 0; it < RS_ISECT_BLOCK / RS_ISECT_THREADS; ++it) { // (uniform trip count: the warp version syncs)
        const int64_t idx = base + it * RS_ISECT_THREADS + threadIdx.x;
        const bool in = idx < a.n_elems;
        int2 r = make_int2(0, 0);
        float2 m = make_float2(0.f, 0.f);
        if (in) {
            r = reinterpret_cast<const int2 *>(a.radii)[idx];
            if (r.x > 0 && r.y > 0)
                m = reinterpret_cast<const float2 *>(a.means2d)[idx];
        }
        const bool vis = in && r.x > 0 && r.y > 0;
        uint4 fp = make_uint4(0u, 0u, 0u, 0u);
        int cnt = 0;
        if (conics != nullptr) {
            float ca = 0.f, cb = 0.f, cc = 0.f, op = 0.f;
            if (vis) {
                ca = conics[idx * 3 + 0];
                cb = conics[idx * 3 + 1];
                cc = conics[idx * 3 + 2];
                op = opacities[idx];
            }
            cnt = rs_tile_footprint_warp(vis, m.x, m.y, r.x, r.y, ca, cb, cc, op, (uint32_t)a.tile_size,
                                         (uint32_t)a.tile_width, (uint32_t)a.tile_height, fp, foot[threadIdx.x >> 5]);
        } else if (vis) {
            const RsTileRect tr = rs_tile_rect(m.x, m.y, (float)r.x, (float)r.y, (uint32_t)a.tile_size,
                                               (uint32_t)a.tile_width, (uint32_t)a.tile_height);
            const uint32_t w = tr.x1 - tr.x0, h = tr.y1 - tr.y0, n = w * h;
            cnt = (int)n;
            if (n > 0u) {
                const unsigned long long mask = n >= 64u ? ~0ull : ((1ull << n) - 1ull);
                fp = make_uint4((uint32_t)mask, (uint32_t)(mask >> 32), tr.x0 | (tr.y0 << 16), w | (h << 16));
            }
        }
        if (in) {
            a.tiles_per_gauss[idx] = cnt;
            footprints[idx] = fp;
            mine += cnt;
        }
    }
    if (a.block_sums != nullptr) { // optional: the per-block sums rs_isect_scan turns into offsets and the total
        const int s = rs_block_sum_256(mine, sums);
        if (threadIdx.x == 0)
            a.block_sums[blockIdx.x] = s;
    }
}

// Exclusive scan of `nb` block sums (in place) by a single CTA; block_sums[nb] and *n_isects receive the total.
// nb is small (1M elements -> 977 blocks), so a serial-over-chunks CTA scan is launch-latency bound.
#define RS_SCAN_THREADS 1024
__global__ void __launch_bounds__(RS_SCAN_THREADS)
rs_isect_scan_kernel(int32_t *block_sums, int nb, int32_t *n_isects, int64_t capacity, int32_t *overflow) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < nb; start += RS_SCAN_THREADS) {
        const int i = start + threadIdx.x;
        long long v = (i < nb) ? block_sums[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += n;
        }
        if (lane == 31)
            warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o)
                    wi += n;
            }
            warp_tot[lane] = wi - w; // exclusive prefix of warp totals
        }
        __syncthreads();
        const long long carry = carry_s;
        const long long excl = carry + warp_tot[warp] + incl - v;
        if (i < nb)
            block_sums[i] = (int32_t)min(excl, (long long)INT32_MAX);
        __syncthreads();
        if (threadIdx.x == RS_SCAN_THREADS - 1)
            carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long total = carry_s;
        block_sums[nb] = (int32_t)min(total, (long long)INT32_MAX);
        if (n_isects != nullptr)
            *n_isects = (int32_t)min(total, (long long)INT32_MAX);
        if (overflow != nullptr)
            *overflow = (total > capacity || total > (long long)INT32_MAX) ? 1 : 0;
    }
}

struct EmitSmem {
    int32_t excl[RS_ISECT_BLOCK + 1]; // exclusive scan of this block's tile counts
    uint32_t rect[RS_ISECT_BLOCK];    // x0 | y0 << 16
    uint32_t width[RS_ISECT_BLOCK];   // x1 - x0
    uint32_t depth[RS_ISECT_BLOCK];   // raw float bits
    int32_t warp_tot[8];
};

__global__ void __launch_bounds__(RS_ISECT_THREADS) rs_isect_emit_kernel(const rs_isect_args a, uint32_t tile_n_bits) {
    __shared__ EmitSmem sm;
    const int64_t base = (int64_t)blockIdx.x * RS_ISECT_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // each thread owns 4 CONSECUTIVE elements (blocked arrangement) so the scan is a per-thread serial prefix
    // followed by one CTA scan of 256 thread totals.
    int cnt[4];
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = threadIdx.x * 4 + k;
        const int64_t idx = base + e;
        int c = 0;
        if (idx < a.n_elems) {
            c = a.tiles_per_gauss[idx];
            if (c > 0) {
                const int2 r = reinterpret_cast<const int2 *>(a.radii)[idx];
                const float2 m = reinterpret_cast<const float2 *>(a.means2d)[idx];
                const RsTileRect tr = rs_tile_rect(m.x, m.y, (float)r.x, (float)r.y, (uint32_t)a.tile_size,
                                                   (uint32_t)a.tile_width, (uint32_t)a.tile_height);
                sm.rect[e] = tr.x0 | (tr.y0 << 16);
                sm.width[e] = tr.x1 - tr.x0;
                sm.depth[e] = __float_as_uint(a.depths[idx]);
            }
        }
        cnt[k] = c;
        tsum += c;
    }
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += n;
    }
    if (lane == 31)
        sm.warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        wbase += (w < warp) ? sm.warp_tot[w] : 0;
    int run = wbase + incl - tsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sm.excl[threadIdx.x * 4 + k] = run;
        run += cnt[k];
    }
    if (threadIdx.x == RS_ISECT_THREADS - 1)
        sm.excl[RS_ISECT_BLOCK] = run;
    __syncthreads();

    const int total = sm.excl[RS_ISECT_BLOCK];
    const int64_t out_base = a.block_sums[blockIdx.x]; // exclusive offset of this block (after rs_isect_scan)
    for (int j = threadIdx.x; j < total; j += RS_ISECT_THREADS) {
        // largest e with excl[e] <= j   (counts of zero are skipped automatically)
        int lo = 0, hi = RS_ISECT_BLOCK;
#pragma unroll
        for (int step = 0; step < 10; ++step) { // log2(1024)
            const int mid = (lo + hi) >> 1;
            if (sm.excl[mid] <= j)
                lo = mid;
            else
                hi = mid;
        }
        const int e = lo;
        const uint32_t r = (uint32_t)(j - sm.excl[e]);
        const uint32_t w = sm.width[e];
        const uint32_t ty = (sm.rect[e] >> 16) + r / w;
        const uint32_t tx = (sm.rect[e] & 0xffffu) + r % w;
        const int64_t idx = base + e;
        const int64_t iid = (a.image_ids != nullptr) ? a.image_ids[idx] : (idx / a.N);
        const int64_t tile_id = (int64_t)ty * a.tile_width + tx;
        const int64_t key = (iid << (32 + tile_n_bits)) | (tile_id << 32) | (int64_t)sm.depth[e];
        const int64_t o = out_base + j;
        if (o < a.capacity) {
            a.isect_ids[o] = key;
            a.flatten_ids[o] = (int32_t)idx;
        }
    }
}

// csrc/IntersectTile.cu:209-257.  KeyT = int64_t: full isect ids (image | tile | depth); KeyT = uint32_t: the
// (image | tile) half alone, as kept by the depth-ordered binning path below.
template <typename KeyT>
__global__ void __launch_bounds__(256)
rs_isect_offsets_kernel(const KeyT *__restrict__ isect_ids, int64_t n_bound, const int32_t *__restrict__ n_dev,
                        uint32_t I, uint32_t n_tiles, uint32_t tile_n_bits, int32_t *__restrict__ offsets) {
    constexpr int HI = sizeof(KeyT) == 8 ? 32 : 0;
    const int64_t n_isects = (n_dev != nullptr) ? min((int64_t)*n_dev, n_bound) : n_bound;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_isects == 0) { // Intersect.cpp:273-276: offsets.fill_(0)
        for (int64_t i = first; i < (int64_t)I * n_tiles; i += stride)
            offsets[i] = 0;
        return;
    }
    // grid-stride: the grid is sized for the SM count, not for n_bound (which may be a loose capacity); 4 independent
    // (current, previous) load pairs per thread and iteration
    constexpr int UNROLL = 4;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * UNROLL; base < n_isects; base += stride * UNROLL) {
        int64_t cur[UNROLL], prev[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t idx = base + (int64_t)u * blockDim.x + threadIdx.x;
            cur[u] = idx < n_isects ? (int64_t)(isect_ids[idx] >> HI) : 0;
            prev[u] = (idx < n_isects && idx > 0) ? (int64_t)(isect_ids[idx - 1] >> HI) : 0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t idx = base + (int64_t)u * blockDim.x + threadIdx.x;
            if (idx >= n_isects)
                continue;
            // only the few thousand boundary elements do any index arithmetic; every other element is two compares
            const bool boundary = idx > 0 && prev[u] != cur[u];
            if (!(boundary || idx == 0 || idx == n_isects - 1))
                continue;
            const int64_t id_curr = (cur[u] >> tile_n_bits) * n_tiles + (cur[u] & ((1ll << tile_n_bits) - 1));
            if (idx == 0) {
                for (int64_t i = 0; i < id_curr + 1; ++i)
                    offsets[i] = 0;
            }
            if (idx == n_isects - 1) {
                for (int64_t i = id_curr + 1; i < (int64_t)I * n_tiles; ++i)
                    offsets[i] = (int32_t)n_isects;
            }
            if (boundary) {
                const int64_t id_prev = (prev[u] >> tile_n_bits) * n_tiles + (prev[u] & ((1ll << tile_n_bits) - 1));
                for (int64_t i = id_prev + 1; i < id_curr + 1; ++i)
                    offsets[i] = (int32_t)idx;
            }
        }
    }
}

// The same table from the 32-bit (image | tile) keys the depth-ordered binning keeps, 4 keys per thread and iteration: one
// 128-bit load, the predecessor of the first key from the neighbouring lane (one extra load only in lane 0), and index
// arithmetic only at the few thousand boundaries.  ~1/5 of the instructions of the generic kernel above.
__global__ void __launch_bounds__(256)
rs_isect_offsets32_kernel(const uint32_t *__restrict__ keys, int64_t n_bound, const int32_t *__restrict__ n_dev, uint32_t I,
                          uint32_t n_tiles, uint32_t tile_n_bits, int32_t *__restrict__ offsets) {
    const int64_t n = (n_dev != nullptr) ? min((int64_t)*n_dev, n_bound) : n_bound;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total_tiles = (int64_t)I * n_tiles;
    if (n == 0) { // Intersect.cpp:273-276: offsets.fill_(0)
        for (int64_t i = first; i < total_tiles; i += stride)
            offsets[i] = 0;
        return;
    }
    const uint32_t tmask = (1u << tile_n_bits) - 1u;
    const int lane = threadIdx.x & 31;
    const int64_t n4 = (n + 3) >> 2; // groups of 4 keys (the buffer is 16-byte aligned; a ragged tail is masked below)
    for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x; q0 < n4; q0 += stride) { // warp-uniform trip count
        const int64_t q = q0 + threadIdx.x;
        const bool live = q < n4;
        uint4 k4 = make_uint4(0u, 0u, 0u, 0u);
        if (live) {
            if (q * 4 + 3 < n) {
                k4 = reinterpret_cast<const uint4 *>(keys)[q];
            } else {
                k4.x = keys[q * 4];
                k4.y = (q * 4 + 1 < n) ? keys[q * 4 + 1] : k4.x;
                k4.z = (q * 4 + 2 < n) ? keys[q * 4 + 2] : k4.y;
                k4.w = k4.z;
            }
        }
        uint32_t prev = __shfl_up_sync(0xffffffffu, k4.w, 1);
        if (lane == 0 && live && q > 0)
            prev = keys[q * 4 - 1];
        if (!live)
            continue;
        const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t idx = q * 4 + u;
            if (idx >= n)
                break;
            const uint32_t cur = kk[u];
            const bool boundary = idx > 0 && cur != prev;
            if (boundary || idx == 0 || idx == n - 1) {
                const int64_t id_curr = (int64_t)(cur >> tile_n_bits) * n_tiles + (cur & tmask);
                if (idx == 0)
                    for (int64_t i = 0; i <= id_curr; ++i)
                        offsets[i] = 0;
                if (idx == n - 1)
                    for (int64_t i = id_curr + 1; i < total_tiles; ++i)
                        offsets[i] = (int32_t)n;
                if (boundary) {
                    const int64_t id_prev = (int64_t)(prev >> tile_n_bits) * n_tiles + (prev & tmask);
                    for (int64_t i = id_prev + 1; i <= id_curr; ++i)
                        offsets[i] = (int32_t)idx;
                }
            }
            prev = cur;
        }
    }
}

static int check_isect_args(const rs_isect_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->n_elems >= 0 && a->I >= 0, "%s: negative sizes", who);
    RS_CHECK(a->tile_size > 0 && a->tile_width > 0 && a->tile_height > 0, "%s: bad tile geometry", who);
    RS_CHECK(a->tile_width < 65536 && a->tile_height < 65536, "%s: tile grid too large", who);
    // Intersect.cpp:50-54: image and tile ids share the upper 32 key bits
    RS_CHECK(rs_bit_width((uint32_t)a->I) + rs_bit_width((uint32_t)(a->tile_width * a->tile_height)) <= 32,
             "%s: image_n_bits + tile_n_bits > 32", who);
    return 0;
}

extern "C" int rs_isect_count(const rs_isect_args *a, rs_stream_t stream) {
    if (int e = check_isect_args(a, "rs_isect_count"))
        return e;
    if (a->n_elems == 0)
        return 0;
    RS_CHECK(a->means2d && a->radii && a->tiles_per_gauss && a->block_sums, "rs_isect_count: null pointer");
    rs_isect_count_kernel<<<rs_isect_num_blocks(a->n_elems), RS_ISECT_THREADS, 0, (cudaStream_t)stream>>>(*a);
    RS_LAUNCH_CHECK("rs_isect_count_kernel");
    return 0;
}

extern "C" int rs_isect_footprints(const rs_isect_args *a, const float *conics, const float *opacities,
                                   uint32_t *tile_footprints, rs_stream_t stream) {
    if (int e = check_isect_args(a, "rs_isect_footprints"))
        return e;
    if (a->n_elems == 0)
        return 0;
    RS_CHECK(a->means2d && a->radii && a->tiles_per_gauss && tile_footprints, "rs_isect_footprints: null pointer");
    RS_CHECK((conics == nullptr) == (opacities == nullptr), "rs_isect_footprints: conics and opacities go together");
    RS_CHECK(a->tile_width < 65536 && a->tile_height < 65536, "rs_isect_footprints: more than 65535 tiles per axis");
    RS_CHECK((reinterpret_cast<uintptr_t>(tile_footprints) & 15) == 0, "rs_isect_footprints: tile_footprints must be 16-byte aligned");
    rs_isect_footprints_kernel<<<rs_isect_num_blocks(a->n_elems), RS_ISECT_THREADS, 0, (cudaStream_t)stream>>>(
        *a, conics, opacities, reinterpret_cast<uint4 *>(tile_footprints));
    RS_LAUNCH_CHECK("rs_isect_footprints_kernel");
    return 0;
}

extern "C" int rs_isect_scan(const rs_isect_args *a, rs_stream_t stream) {
    if (int e = check_isect_args(a, "rs_isect_scan"))
        return e;
    RS_CHECK(a->block_sums && a->n_isects, "rs_isect_scan: null pointer");
    rs_isect_scan_kernel<<<1, RS_SCAN_THREADS, 0, (cudaStream_t)stream>>>(
        a->block_sums, rs_isect_num_blocks(a->n_elems), a->n_isects, a->capacity > 0 ? a->capacity : INT64_MAX,
        a->overflow);
    RS_LAUNCH_CHECK("rs_isect_scan_kernel");
    return 0;
}

extern "C" int rs_isect_count_total(const rs_isect_args *a, rs_stream_t stream, int64_t *n_isects_host) {
    RS_CHECK(a && a->n_isects && n_isects_host, "rs_isect_count_total: null pointer");
    int32_t h = 0;
    RS_CUDA(cudaMemcpyAsync(&h, a->n_isects, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    RS_CHECK(h != INT32_MAX, "rs_isect_count_total: more than 2^31-1 tile intersections");
    *n_isects_host = h;
    return 0;
}

extern "C" int rs_isect_emit(const rs_isect_args *a, rs_stream_t stream) {
    if (int e = check_isect_args(a, "rs_isect_emit"))
        return e;
    if (a->n_elems == 0)
        return 0;
    RS_CHECK(a->means2d && a->radii && a->depths && a->tiles_per_gauss && a->block_sums && a->isect_ids &&
                 a->flatten_ids,
             "rs_isect_emit: null pointer");
    RS_CHECK(a->image_ids != nullptr || a->N > 0, "rs_isect_emit: N required when not packed");
    const uint32_t tile_n_bits = rs_bit_width((uint32_t)(a->tile_width * a->tile_height));
    rs_isect_emit_kernel<<<rs_isect_num_blocks(a->n_elems), RS_ISECT_THREADS, 0, (cudaStream_t)stream>>>(*a,
                                                                                                          tile_n_bits);
    RS_LAUNCH_CHECK("rs_isect_emit_kernel");
    return 0;
}

extern "C" int rs_isect_offsets(const int64_t *isect_ids_sorted, int64_t n_isects, const int32_t *n_isects_dev,
                                int32_t I, int32_t tile_width, int32_t tile_height, int32_t *offsets,
                                rs_stream_t stream) {
    RS_CHECK(offsets != nullptr, "rs_isect_offsets: null offsets");
    RS_CHECK(I >= 0 && tile_width > 0 && tile_height > 0, "rs_isect_offsets: bad geometry");
    if ((int64_t)I * tile_width * tile_height == 0)
        return 0;
    RS_CHECK(n_isects == 0 || isect_ids_sorted != nullptr, "rs_isect_offsets: null isect_ids");
    const uint32_t n_tiles = (uint32_t)(tile_width * tile_height);
    const uint32_t tile_n_bits = rs_bit_width(n_tiles);
    // grid-stride kernel: at most 8 CTAs per SM; at least one CTA per SM so that a device-side count of zero still
    // clears the table quickly
    int64_t grid = (n_isects + 255) / 256;
    grid = min(grid, (int64_t)rs_num_sms() * 8);
    if (grid < 1 || n_isects_dev != nullptr)
        grid = max(grid, (int64_t)rs_num_sms());
    rs_isect_offsets_kernel<int64_t><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        isect_ids_sorted, n_isects, n_isects_dev, (uint32_t)I, n_tiles, tile_n_bits, offsets);
    RS_LAUNCH_CHECK("rs_isect_offsets_kernel");
    return 0;
}

// =====================================================================================================================
// Depth-ordered binning: rs_isect_sorted.
//
// The reference sorts M 64-bit (image | tile | depth) keys with cub (csrc/IntersectTile.cu:296-339): ceil(46/8) = 6 LSD
// passes over 12 B pairs for a 1080p frame.  The same order -- (image, tile, depth bits, flatten index), stable --
// is produced here with far less traffic by sorting the two halves of the key separately:
//   1. the V visible elements in (depth bits, flatten index) order: bucket sort (depth_order.cu)   ~20 B per element
//   2. count / scan / emit in that depth order: (image | tile) 32-bit keys, flatten-index values
//   3. stable LSD sort of the M (image | tile, flatten index) pairs on the tile + image bits only
//                                                                                 ceil(14/8) = 2 passes over M 8-byte pairs
//   4. offsets from the sorted 32-bit keys; the 64-bit isect ids are rebuilt (tile key << 32 | depth bits of the
//      flatten id) only when the caller asks for them.
// Ties: equal depths are listed in ascending flatten index by 1, every Gaussian contributes at most one
// intersection per tile, and 3 is stable, so equal (image, tile, depth) keys stay in ascending flatten-index order --
// exactly the order the reference's stable sort of its emission order gives.
// =====================================================================================================================
int rs_sort_pairs_u32_internal(int64_t n_bound, const int32_t *n_dev, int begin_bit, int end_bit, const uint32_t *keys_in,
                               const int32_t *vals_in, uint32_t *kbuf0, int32_t *vbuf0, uint32_t *kbuf1, int32_t *vbuf1,
                               void *workspace, uint64_t workspace_bytes, int *passes, cudaStream_t s, bool hist_ready);
int rs_sort_ws_prepare(void *workspace, cudaStream_t s);
// depth_order.cu
uint64_t rs_depth_order_workspace_bytes(int64_t n_elems);
int rs_depth_order(int64_t n_elems, const float *depths, const int32_t *tiles, int32_t *elems_out, int32_t *n_sorted_dev,
                   void *workspace, uint64_t workspace_bytes, cudaStream_t s, bool stats_ready);
int rs_depth_order_prepare(void *workspace, int64_t n_elems, cudaStream_t s);
uint32_t *rs_depth_order_stats_ptr(void *workspace, int64_t n_elems);

// block sums of the tile counts taken in depth order
__global__ void __launch_bounds__(RS_ISECT_THREADS)
rs_bin_count_kernel(int64_t n_bound, const int32_t *__restrict__ n_sorted, const int32_t *__restrict__ elems,
                    const int32_t *__restrict__ tiles_per_gauss, int32_t *__restrict__ block_sums) {
    __shared__ int sums[8];
    const int64_t n_elems = min((int64_t)*n_sorted, n_bound); // visible elements, in depth order
    const int64_t base = (int64_t)blockIdx.x * RS_ISECT_BLOCK;
    int mine = 0;
#pragma unroll
    for (int it = 0; it < RS_ISECT_BLOCK / RS_ISECT_THREADS; ++it) {
        const int64_t i = base + it * RS_ISECT_THREADS + threadIdx.x;
        if (i < n_elems)
            mine += tiles_per_gauss[elems[i]];
    }
    const int s = rs_block_sum_256(mine, sums);
    if (threadIdx.x == 0)
        block_sums[blockIdx.x] = s;
}

// Emission in depth order.  A CTA owns RS_ISECT_BLOCK depth-consecutive elements (4 per thread, blocked); after a CTA
// scan of their tile counts, every thread walks the tiles of its own elements (row-major, no division in the loop) and
// writes (image|tile key, flatten id) into a shared-memory window of EMIT_CHUNK output slots; the window is then copied
// to global memory with fully coalesced stores.  Elements covering more than EMIT_BIG tiles (a splat filling the
// screen) are filled by the whole CTA instead of one thread, so a single huge Gaussian cannot serialise the CTA.
#define EMIT_CHUNK 4096
#define EMIT_BIG 64
struct BinEmitSmem {
    int32_t excl[RS_ISECT_BLOCK + 1];
    uint32_t skey[EMIT_CHUNK];
    int32_t sval[EMIT_CHUNK];
    // elements with more than EMIT_BIG tiles
    int32_t big_start[RS_ISECT_BLOCK / 4];
    uint32_t big_rect[RS_ISECT_BLOCK / 4]; // x0 | y0 << 16
    uint32_t big_wh[RS_ISECT_BLOCK / 4];   // width | count... (width in the low 16 bits)
    int32_t big_cnt[RS_ISECT_BLOCK / 4];
    uint32_t big_hi[RS_ISECT_BLOCK / 4];
    int32_t big_elem[RS_ISECT_BLOCK / 4];
    int32_t n_big;
    int32_t warp_tot[8];
    unsigned int hist[4 * 256]; // up to 4 passes of 8-bit digits over the (image | tile) bits
};

// FOOTPRINTS: the tiles of an element come from its tile footprint (rs_project_fwd_args.tile_footprints, one 16-byte load
// instead of radii + means2d + count): bounding rectangle + a 64-bit mask of the listed tiles when the rectangle has at
// most 64 of them.
template <bool FOOTPRINTS>
__global__ void __launch_bounds__(RS_ISECT_THREADS)
rs_bin_emit_kernel(const rs_isect_args a, const uint4 *__restrict__ footprints, const int32_t *__restrict__ elems, const int32_t *__restrict__ n_sorted,
                   uint32_t tile_n_bits, uint32_t *__restrict__ tile_keys, int32_t *__restrict__ vals,
                   uint32_t *__restrict__ sort_ws, int sort_bits, int sort_nb_stride) {
    constexpr bool TIGHT = FOOTPRINTS;
    __shared__ BinEmitSmem sm;
    const int64_t n_live = min((int64_t)*n_sorted, (int64_t)a.n_elems); // visible elements, in depth order
    // digit histograms of the keys this CTA emits, for every pass of the tile sort that follows (saves that sort its own
    // read of all M keys)
    const int sort_passes = sort_num_passes(sort_bits), sort_width = sort_digit_width(sort_bits);
    for (int i = threadIdx.x; i < sort_passes * RADIX; i += RS_ISECT_THREADS)
        sm.hist[i] = 0;
    const int64_t base = (int64_t)blockIdx.x * RS_ISECT_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        sm.n_big = 0;
    int cnt[4], elem[4];
    uint32_t x0[4], y0[4], w[4], hi[4];
    unsigned long long mask[TIGHT ? 4 : 1];
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = base + threadIdx.x * 4 + k;
        int c = 0;
        elem[k] = 0;
        x0[k] = y0[k] = hi[k] = 0;
        w[k] = 1;
        if (TIGHT)
            mask[TIGHT ? k : 0] = 0ull;
        if (TIGHT) {
            if (i < n_live) {
                const int32_t e = elems[i];
                const uint4 f = footprints[e];
                const uint32_t fw = f.w & 0xffffu, fh = f.w >> 16, n = fw * fh;
                const unsigned long long m = (unsigned long long)f.x | ((unsigned long long)f.y << 32);
                c = (n <= 64u) ? __popcll(m) : (int)n;
                if (c > 0) {
                    x0[k] = f.z & 0xffffu;
                    y0[k] = f.z >> 16;
                    w[k] = fw;
                    mask[TIGHT ? k : 0] = m;
                    const uint32_t iid = (a.image_ids != nullptr) ? (uint32_t)a.image_ids[e] : (uint32_t)(e / a.N);
                    hi[k] = iid << tile_n_bits;
                    elem[k] = e;
                }
            }
        } else if (i < n_live) {
            const int32_t e = elems[i];
            c = a.tiles_per_gauss[e];
            if (c > 0) {
                const int2 r = reinterpret_cast<const int2 *>(a.radii)[e];
                const float2 m = reinterpret_cast<const float2 *>(a.means2d)[e];
                const RsTileRect tr = rs_tile_rect(m.x, m.y, (float)r.x, (float)r.y, (uint32_t)a.tile_size,
                                                   (uint32_t)a.tile_width, (uint32_t)a.tile_height);
                x0[k] = tr.x0;
                y0[k] = tr.y0;
                w[k] = tr.x1 - tr.x0;
                const uint32_t iid = (a.image_ids != nullptr) ? (uint32_t)a.image_ids[e] : (uint32_t)(e / a.N);
                hi[k] = iid << tile_n_bits;
                elem[k] = e;
            }
        }
        cnt[k] = c;
        tsum += c;
    }
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += n;
    }
    if (lane == 31)
        sm.warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int t = sm.warp_tot[q];
        wbase += (q < warp) ? t : 0;
        total += t;
    }
    int start[4];
    {
        int run = wbase + incl - tsum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            start[k] = run;
            run += cnt[k];
            if (cnt[k] > EMIT_BIG) { // handed to the whole CTA
                const int slot = atomicAdd(&sm.n_big, 1);
                if (slot < RS_ISECT_BLOCK / 4) {
                    sm.big_start[slot] = start[k];
                    sm.big_rect[slot] = x0[k] | (y0[k] << 16);
                    sm.big_wh[slot] = w[k];
                    sm.big_cnt[slot] = cnt[k];
                    sm.big_hi[slot] = hi[k];
                    sm.big_elem[slot] = elem[k];
                    cnt[k] = 0; // not filled by this thread
                }
            }
        }
    }
    __syncthreads();
    const int n_big = min(sm.n_big, RS_ISECT_BLOCK / 4);
    const int64_t out_base = a.block_sums[blockIdx.x]; // exclusive offset of this block (after the scan kernel)
    const uint32_t tile_w = (uint32_t)a.tile_width;

    for (int B = 0; B < total; B += EMIT_CHUNK) {
        const int Bend = B + EMIT_CHUNK;
        // every thread: the tiles of its own (small) elements that fall into this window
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int s0 = start[k], s1 = start[k] + cnt[k];
            const int lo = max(s0, B), up = min(s1, Bend);
            if (TIGHT && cnt[k] <= 64 && lo < up) { // masked rectangles (more than 64 tiles: all of them, below)
                unsigned long long mm = mask[TIGHT ? k : 0];
                for (int skip = lo - s0; skip > 0; --skip)
                    mm &= mm - 1ull; // listed tiles that went into earlier windows
                const uint32_t inv = 65536u / w[k] + 1u; // t / w for t < 64, w <= 64
                for (int sl = lo; sl < up; ++sl) {
                    const uint32_t t = (uint32_t)__ffsll((long long)mm) - 1u;
                    mm &= mm - 1ull;
                    const uint32_t row = (t * inv) >> 16;
                    sm.skey[sl - B] = hi[k] | ((y0[k] + row) * tile_w + x0[k] + (t - row * w[k]));
                    sm.sval[sl - B] = elem[k];
                }
            } else if (lo < up) {
                const uint32_t r = (uint32_t)(lo - s0);
                uint32_t ty = y0[k] + r / w[k];
                uint32_t tx = x0[k] + r % w[k];
                const uint32_t xe = x0[k] + w[k];
                for (int sl = lo; sl < up; ++sl) {
                    sm.skey[sl - B] = hi[k] | (ty * tile_w + tx);
                    sm.sval[sl - B] = elem[k];
                    if (++tx == xe) {
                        tx = x0[k];
                        ++ty;
                    }
                }
            }
        }
        // the whole CTA: big elements
        for (int q = 0; q < n_big; ++q) {
            const int s0 = sm.big_start[q], s1 = s0 + sm.big_cnt[q];
            const int lo = max(s0, B), up = min(s1, Bend);
            const uint32_t bw = sm.big_wh[q], bx0 = sm.big_rect[q] & 0xffffu, by0 = sm.big_rect[q] >> 16;
            for (int sl = lo + threadIdx.x; sl < up; sl += RS_ISECT_THREADS) {
                const uint32_t r = (uint32_t)(sl - s0);
                sm.skey[sl - B] = sm.big_hi[q] | ((by0 + r / bw) * tile_w + bx0 + r % bw);
                sm.sval[sl - B] = sm.big_elem[q];
            }
        }
        __syncthreads();
        const int n_out = min(EMIT_CHUNK, total - B);
        for (int j0 = 0; j0 < n_out; j0 += RS_ISECT_THREADS) { // warp-uniform trip count (ballots below)
            const int jj = j0 + threadIdx.x;
            const int64_t o = out_base + B + jj;
            const bool live = jj < n_out && o < a.capacity;
            const uint32_t key = live ? sm.skey[jj] : 0u;
            if (live) {
                tile_keys[o] = key;
                vals[o] = sm.sval[jj];
            }
            const unsigned act = __ballot_sync(0xffffffffu, live);
            if (act != 0u) {
                for (int p = 0; p < sort_passes; ++p) {
                    const int bits = min(sort_width, sort_bits - p * sort_width);
                    const uint32_t d = (key >> (p * sort_width)) & ((1u << bits) - 1u);
                    const uint32_t d0 = __shfl_sync(0xffffffffu, d, __ffs(act) - 1);
                    const unsigned same = __ballot_sync(0xffffffffu, live && d == d0);
                    if (same == act) { // a whole warp on one digit (the high tile bits): one add
                        if (lane == __ffs(act) - 1)
                            atomicAdd(&sm.hist[p * RADIX + d0], (unsigned)__popc(act));
                    } else if (live) {
                        atomicAdd(&sm.hist[p * RADIX + d], 1u);
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < sort_passes * RADIX; i += RS_ISECT_THREADS) {
        const unsigned v = sm.hist[i];
        if (v)
            atomicAdd(&sort_ws[WS_HIST + i], v);
    }
    // clear the look-back words the tile sort will use (the device-side count is known since the scan kernel)
    {
        const int64_t n = min((int64_t)*a.n_isects, a.capacity);
        const int64_t words = (n + SORT_TILE - 1) / SORT_TILE * RADIX;
        for (int p = 0; p < sort_passes; ++p) {
            uint32_t *lb = sort_ws + WS_LOOKBACK + (size_t)p * sort_nb_stride * RADIX;
            for (int64_t i = (int64_t)blockIdx.x * RS_ISECT_THREADS + threadIdx.x; i < words;
                 i += (int64_t)gridDim.x * RS_ISECT_THREADS)
                lb[i] = 0u;
        }
    }
}

// isect_ids[i] = tile_key[i] << 32 | depth bits of flatten id i (csrc/IntersectTile.cu:95-108)
__global__ void __launch_bounds__(256)
rs_bin_keys64_kernel(int64_t n_bound, const int32_t *__restrict__ n_dev, const uint32_t *__restrict__ tile_keys,
                     const int32_t *__restrict__ vals, const float *__restrict__ depths, int64_t *__restrict__ isect_ids) {
    const int64_t n = (n_dev != nullptr) ? min((int64_t)*n_dev, n_bound) : n_bound;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        isect_ids[i] = ((int64_t)tile_keys[i] << 32) | (int64_t)__float_as_uint(depths[vals[i]]);
}

namespace {
struct BinLayout {
    size_t elems, n_sorted, dord_ws, dord_ws_bytes, block_sums, tkeys_a, tkeys_b, vals_b, sort_ws, sort_ws_bytes, total;
};
inline size_t bin_align(size_t x) { return (x + 255) & ~(size_t)255; }
BinLayout bin_layout(int64_t n_elems, int64_t capacity) {
    BinLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = bin_align(o + bytes);
        return at;
    };
    const size_t E = (size_t)(n_elems > 0 ? n_elems : 1), M = (size_t)(capacity > 0 ? capacity : 1);
    L.elems = take(E * 4);
    L.n_sorted = take(4);
    L.dord_ws_bytes = rs_depth_order_workspace_bytes((int64_t)E);
    L.dord_ws = take(L.dord_ws_bytes);
    L.block_sums = take(((size_t)rs_isect_num_blocks((int64_t)E) + 1) * 4);
    L.tkeys_a = take(M * 4);
    L.tkeys_b = take(M * 4);
    L.vals_b = take(M * 4);
    L.sort_ws_bytes = rs_radix_sort_workspace_bytes((int64_t)M);
    L.sort_ws = take(L.sort_ws_bytes);
    L.total = o;
    return L;
}
} // namespace

extern "C" uint64_t rs_isect_sorted_workspace_bytes(int64_t n_elems, int64_t capacity) {
    if (n_elems < 0 || capacity < 0)
        return 0;
    return bin_layout(n_elems, capacity).total;
}

extern "C" int rs_isect_sorted_prepare(void *workspace, int64_t n_elems, int64_t capacity, rs_stream_t stream) {
    RS_CHECK(workspace != nullptr && n_elems >= 0 && capacity >= 0, "rs_isect_sorted_prepare: bad arguments");
    const BinLayout L = bin_layout(n_elems, capacity);
    return rs_depth_order_prepare(reinterpret_cast<char *>(workspace) + L.dord_ws, n_elems > 0 ? n_elems : 1, (cudaStream_t)stream);
}
extern "C" uint32_t *rs_isect_sorted_depth_stats(void *workspace, int64_t n_elems, int64_t capacity) {
    if (workspace == nullptr || n_elems < 0 || capacity < 0)
        return nullptr;
    const BinLayout L = bin_layout(n_elems, capacity);
    return rs_depth_order_stats_ptr(reinterpret_cast<char *>(workspace) + L.dord_ws, n_elems > 0 ? n_elems : 1);
}

extern "C" int rs_isect_sorted(const rs_isect_sorted_args *b, rs_stream_t stream) {
    RS_CHECK(b != nullptr, "rs_isect_sorted: null args");
    const rs_isect_args *a = &b->isect;
    if (int e = check_isect_args(a, "rs_isect_sorted"))
        return e;
    RS_CHECK(a->n_elems < ((int64_t)1 << 30), "rs_isect_sorted: too many elements (limit 2^30)");
    RS_CHECK(a->capacity >= 0 && a->capacity < ((int64_t)1 << 31), "rs_isect_sorted: bad capacity");
    RS_CHECK(a->n_isects != nullptr, "rs_isect_sorted: n_isects (device) required");
    RS_CHECK(a->n_elems == 0 || (a->means2d && a->radii && a->depths && a->tiles_per_gauss),
             "rs_isect_sorted: null input pointer");
    RS_CHECK(a->capacity == 0 || a->flatten_ids != nullptr, "rs_isect_sorted: null flatten_ids");
    RS_CHECK(a->image_ids != nullptr || a->N > 0 || a->n_elems == 0, "rs_isect_sorted: N required when not packed");
    const BinLayout L = bin_layout(a->n_elems, a->capacity);
    RS_CHECK(b->workspace != nullptr && b->workspace_bytes >= L.total, "rs_isect_sorted: workspace too small (%llu < %llu)",
             (unsigned long long)b->workspace_bytes, (unsigned long long)L.total);
    cudaStream_t s = (cudaStream_t)stream;
    char *w = reinterpret_cast<char *>(b->workspace);
    const uint32_t n_tiles = (uint32_t)(a->tile_width * a->tile_height);
    const uint32_t tile_n_bits = rs_bit_width(n_tiles);
    const uint32_t image_n_bits = rs_bit_width((uint32_t)a->I);
    int32_t *block_sums = reinterpret_cast<int32_t *>(w + L.block_sums);
    const int nb = rs_isect_num_blocks(a->n_elems);

    // tile sort: pass p writes buffer p & 1, the last pass must land in the caller's flatten_ids, and the emission goes to
    // "buffer 1" (which the sort allows to alias its input)
    const int tile_bits = (int)(tile_n_bits + image_n_bits);
    const int tile_passes = (tile_bits + 7) / 8;
    uint32_t *tk_final = reinterpret_cast<uint32_t *>(w + L.tkeys_a), *tk_other = reinterpret_cast<uint32_t *>(w + L.tkeys_b);
    int32_t *tv_final = a->flatten_ids, *tv_other = reinterpret_cast<int32_t *>(w + L.vals_b);
    const bool final_is_buf0 = (tile_passes & 1) != 0;
    uint32_t *tk_buf0 = final_is_buf0 ? tk_final : tk_other, *tk_buf1 = final_is_buf0 ? tk_other : tk_final;
    int32_t *tv_buf0 = final_is_buf0 ? tv_final : tv_other, *tv_buf1 = final_is_buf0 ? tv_other : tv_final;
    const uint32_t *tkeys = tk_final;

    const int32_t *elems = reinterpret_cast<const int32_t *>(w + L.elems);
    const int32_t *n_sorted = reinterpret_cast<const int32_t *>(w + L.n_sorted);
    if (a->n_elems > 0) {
        // 1. depth order of the visible elements: bucket sort on the depth bits (depth_order.cu), ties by flatten index
        if (int e = rs_depth_order(a->n_elems, a->depths, a->tiles_per_gauss, reinterpret_cast<int32_t *>(w + L.elems),
                                   reinterpret_cast<int32_t *>(w + L.n_sorted), w + L.dord_ws, L.dord_ws_bytes, s,
                                   b->depth_stats_ready != 0))
            return e;
        // 2a. block sums of the tile counts in depth order
        rs_bin_count_kernel<<<nb, RS_ISECT_THREADS, 0, s>>>(a->n_elems, n_sorted, elems, a->tiles_per_gauss, block_sums);
        RS_LAUNCH_CHECK("rs_bin_count_kernel");
    }
    // 2b. exclusive block offsets, device-side total and overflow flag
    rs_isect_scan_kernel<<<1, RS_SCAN_THREADS, 0, s>>>(block_sums, a->n_elems > 0 ? nb : 0, a->n_isects,
                                                       a->capacity, a->overflow);
    RS_LAUNCH_CHECK("rs_isect_scan_kernel");
    if (a->n_elems > 0 && a->capacity > 0) {
        // 2c. emission in depth order
        rs_isect_args e = *a;
        e.block_sums = block_sums;
        RS_CHECK(tile_passes <= 4, "rs_isect_sorted: image + tile bits exceed 32");
        if (int err = rs_sort_ws_prepare(w + L.sort_ws, s)) // the emission kernel accumulates the sort's histograms
            return err;
        const int sort_nb_stride = (int)((a->capacity + SORT_TILE - 1) / SORT_TILE);
        if (b->tile_footprints != nullptr)
            rs_bin_emit_kernel<true><<<nb, RS_ISECT_THREADS, 0, s>>>(
                e, reinterpret_cast<const uint4 *>(b->tile_footprints), elems, n_sorted, tile_n_bits, tk_buf1, tv_buf1,
                reinterpret_cast<uint32_t *>(w + L.sort_ws), tile_bits, sort_nb_stride);
        else
            rs_bin_emit_kernel<false><<<nb, RS_ISECT_THREADS, 0, s>>>(e, nullptr, elems, n_sorted, tile_n_bits, tk_buf1,
                                                                      tv_buf1, reinterpret_cast<uint32_t *>(w + L.sort_ws),
                                                                      tile_bits, sort_nb_stride);
        RS_LAUNCH_CHECK("rs_bin_emit_kernel");
        // 3. stable sort on the (image | tile) bits only
        int passes = 0;
        if (int err = rs_sort_pairs_u32_internal(a->capacity, a->n_isects, 0, tile_bits, tk_buf1, tv_buf1, tk_buf0, tv_buf0,
                                                 tk_buf1, tv_buf1, w + L.sort_ws, L.sort_ws_bytes, &passes, s, true))
            return err;
    }
    // 4. offsets (+ 64-bit ids on request)
    if (b->tile_offsets != nullptr && (int64_t)a->I * n_tiles > 0) {
        int64_t grid = min((a->capacity + 255) / 256, (int64_t)rs_num_sms() * 8);
        grid = max(grid, (int64_t)rs_num_sms());
        if ((reinterpret_cast<uintptr_t>(tkeys) & 15) == 0)
            rs_isect_offsets32_kernel<<<(unsigned)grid, 256, 0, s>>>(
                tkeys, a->capacity, a->n_isects, (uint32_t)a->I, n_tiles, tile_n_bits, b->tile_offsets);
        else
            rs_isect_offsets_kernel<uint32_t><<<(unsigned)grid, 256, 0, s>>>(tkeys, a->capacity, a->n_isects, (uint32_t)a->I,
                                                                            n_tiles, tile_n_bits, b->tile_offsets);
        RS_LAUNCH_CHECK("rs_isect_offsets_kernel");
    }
    if (a->isect_ids != nullptr && a->capacity > 0) {
        rs_bin_keys64_kernel<<<rs_cdiv(a->capacity, 256), 256, 0, s>>>(a->capacity, a->n_isects, tkeys, a->flatten_ids,
                                                                       a->depths, a->isect_ids);
        RS_LAUNCH_CHECK("rs_bin_keys64_kernel");
    }
    return 0;
}

// Sorted 64-bit isect ids recovered from (offsets, sorted flatten ids, depths): entry i belongs to the last (image, tile)
// whose offset is <= i.  Used to export `meta["isect_ids"]` from the frame path, which never materialises 64-bit keys.
__global__ void __launch_bounds__(256)
rs_isect_ids_from_offsets_kernel(const int32_t *__restrict__ offsets, const int32_t *__restrict__ flatten_ids,
                                 const float *__restrict__ depths, int64_t n_bound, const int32_t *__restrict__ n_dev,
                                 uint32_t n_total_tiles, uint32_t n_tiles, uint32_t tile_n_bits,
                                 int64_t *__restrict__ isect_ids) {
    const int64_t n = (n_dev != nullptr) ? min((int64_t)*n_dev, n_bound) : n_bound;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    uint32_t lo = 0, hi = n_total_tiles; // invariant: offsets[lo] <= i, (hi == n_total_tiles or offsets[hi] > i)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((int64_t)offsets[mid] <= i)
            lo = mid;
        else
            hi = mid;
    }
    const uint32_t image = lo / n_tiles, tile = lo - image * n_tiles;
    const int64_t hi32 = ((int64_t)image << tile_n_bits) | (int64_t)tile;
    isect_ids[i] = (hi32 << 32) | (int64_t)__float_as_uint(depths[flatten_ids[i]]);
}

int rs_isect_ids_from_offsets(const int32_t *offsets, const int32_t *flatten_ids, const float *depths, int64_t n_bound,
                              const int32_t *n_dev, int32_t I, int32_t tile_width, int32_t tile_height,
                              int64_t *isect_ids, rs_stream_t stream) {
    RS_CHECK(offsets && flatten_ids && depths && isect_ids, "rs_isect_ids_from_offsets: null pointer");
    if (n_bound <= 0)
        return 0;
    const uint32_t n_tiles = (uint32_t)(tile_width * tile_height);
    rs_isect_ids_from_offsets_kernel<<<rs_cdiv(n_bound, 256), 256, 0, (cudaStream_t)stream>>>(
        offsets, flatten_ids, depths, n_bound, n_dev, (uint32_t)I * n_tiles, n_tiles, rs_bit_width(n_tiles), isect_ids);
    RS_LAUNCH_CHECK("rs_isect_ids_from_offsets_kernel");
    return 0;
}
