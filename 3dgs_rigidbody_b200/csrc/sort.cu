// Hand-written stable LSD radix sort of (key, int32 value) pairs -- replaces the cub::DeviceRadixSort::SortPairs call
// of csrc/IntersectTile.cu:296-339 (keys = image | tile | depth bits, values = flatten ids).
//
// 8-bit digits, three kernels per pass, no inter-CTA spinning (every kernel is a plain grid):
//   hist       each CTA counts the digits of its 4096-pair tile            -> hist[digit][cta]
//   scan_rows  one CTA per digit: exclusive scan over CTAs (in place)       -> hist[digit][cta], bin_tot[digit]
//   scatter    each CTA re-reads its tile, ranks pairs stably (warp match + per-warp digit counters, warp-striped
//              order), re-orders the tile in shared memory so that equal digits are contiguous, and writes runs to
//              their global positions (coalesced within a run).
// Stability: tile order = (warp, item, lane) = ascending index; ranks preserve it; CTAs are ordered by the row scan.
// Algorithmic HBM bytes per pair per pass: 8 (hist read) + 12 (read) + 12 (write) for 64-bit keys.
// A device-side pair count (`n_dev`) makes the whole sort launchable without knowing n on the host.
#include "common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 16
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_WARPS (SORT_THREADS / 32)
#define RADIX_BITS 8
#define RADIX (1 << RADIX_BITS)

static_assert(RADIX == SORT_THREADS, "one thread per digit in the block-level scans");

__device__ __forceinline__ int64_t sort_count(int64_t n_bound, const int32_t *n_dev) {
    return (n_dev != nullptr) ? min((int64_t)*n_dev, n_bound) : n_bound;
}

// exclusive scan of one int per thread across a 256-thread CTA; `total` gets the sum.  `wsum` = 8 ints of smem.
__device__ __forceinline__ int block_excl_scan_256(int v, int *wsum, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += n;
    }
    __syncthreads(); // protect wsum from a previous use
    if (lane == 31)
        wsum[warp] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
        const int s = wsum[w];
        base += (w < warp) ? s : 0;
        tot += s;
    }
    if (total != nullptr)
        *total = tot;
    return base + incl - v;
}

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const KeyT *__restrict__ keys, int64_t n_bound, const int32_t *__restrict__ n_dev, int shift,
                 uint32_t mask, int32_t *__restrict__ hist, int nblocks) {
    __shared__ int h[SORT_WARPS][RADIX];
    const int64_t n = sort_count(n_bound, n_dev);
    if ((int64_t)blockIdx.x * SORT_TILE >= n)
        return; // the grid is sized for n_bound; rows are only scanned up to ceil(n / SORT_TILE)
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
        h[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t tile_start = (int64_t)blockIdx.x * SORT_TILE;
    if (tile_start < n) {
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const int64_t idx = tile_start + k * SORT_THREADS + threadIdx.x;
            if (idx < n) {
                const uint32_t d = (uint32_t)(keys[idx] >> shift) & mask;
                atomicAdd(&h[warp][d], 1);
            }
        }
    }
    __syncthreads();
    int s = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
        s += h[w][threadIdx.x];
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s;
}

// grid = RADIX CTAs; CTA d scans row d of hist (nblocks entries) in place (exclusive) and writes bin_tot[d].
__global__ void __launch_bounds__(SORT_THREADS)
sort_scan_rows_kernel(int32_t *__restrict__ hist, int nblocks, int64_t n_bound, const int32_t *__restrict__ n_dev,
                      int32_t *__restrict__ bin_tot) {
    __shared__ int wsum[SORT_WARPS];
    int32_t *row = hist + (size_t)blockIdx.x * nblocks;
    const int nb_eff = (int)((sort_count(n_bound, n_dev) + SORT_TILE - 1) / SORT_TILE);
    int carry = 0;
    for (int start = 0; start < nb_eff; start += SORT_THREADS) {
        const int i = start + threadIdx.x;
        const int v = (i < nb_eff) ? row[i] : 0;
        int tot;
        const int ex = block_excl_scan_256(v, wsum, &tot);
        if (i < nb_eff)
            row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0)
        bin_tot[blockIdx.x] = carry;
}

template <typename KeyT> struct SortSmem {
    KeyT keys[SORT_TILE];
    int32_t vals[SORT_TILE];
    int wh[SORT_WARPS][RADIX]; // per-warp digit counters -> exclusive per-warp offsets
    int bin_start[RADIX];      // first slot of each digit inside the re-ordered tile
    int gbase[RADIX];          // global index of that first slot
    int wsum[SORT_WARPS];
};

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const KeyT *__restrict__ keys_in, const int32_t *__restrict__ vals_in, KeyT *__restrict__ keys_out,
                    int32_t *__restrict__ vals_out, int64_t n_bound, const int32_t *__restrict__ n_dev, int shift,
                    uint32_t mask, const int32_t *__restrict__ hist, int nblocks,
                    const int32_t *__restrict__ bin_tot) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem<KeyT> &sm = *reinterpret_cast<SortSmem<KeyT> *>(sort_smem_raw);
    const int64_t n = sort_count(n_bound, n_dev);
    const int64_t tile_start = (int64_t)blockIdx.x * SORT_TILE;
    if (tile_start >= n)
        return;
    const int count = (int)min((int64_t)SORT_TILE, n - tile_start);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = rs_lanemask_lt();

#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
        sm.wh[w][threadIdx.x] = 0;

    // warp-striped load: warp w owns [w*512, (w+1)*512), item k of lane l sits at k*32 + l
    KeyT key[SORT_ITEMS];
    int32_t val[SORT_ITEMS];
    const int wbase = warp * (32 * SORT_ITEMS);
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int local = wbase + k * 32 + lane;
        if (local < count) {
            key[k] = keys_in[tile_start + local];
            val[k] = vals_in[tile_start + local];
        } else {
            key[k] = ~(KeyT)0; // sorts to the very end of the tile, never written out
            val[k] = 0;
        }
    }
    __syncthreads();

    int rank[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const uint32_t d = (uint32_t)(key[k] >> shift) & mask;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        int old = 0;
        if (lane == leader) {
            old = sm.wh[warp][d];
            sm.wh[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // thread t handles digit t: per-warp exclusive offsets, tile-level digit start, global base
    {
        int sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const int c = sm.wh[w][threadIdx.x];
            sm.wh[w][threadIdx.x] = sum;
            sum += c;
        }
        const int start = block_excl_scan_256(sum, sm.wsum, nullptr);
        const int gb = block_excl_scan_256(bin_tot[threadIdx.x], sm.wsum, nullptr);
        sm.bin_start[threadIdx.x] = start;
        sm.gbase[threadIdx.x] = gb + hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const uint32_t d = (uint32_t)(key[k] >> shift) & mask;
        const int pos = sm.bin_start[d] + sm.wh[warp][d] + rank[k];
        sm.keys[pos] = key[k];
        sm.vals[pos] = val[k];
    }
    __syncthreads();

    for (int i = threadIdx.x; i < count; i += SORT_THREADS) {
        const KeyT kk = sm.keys[i];
        const uint32_t d = (uint32_t)(kk >> shift) & mask;
        const int64_t out = (int64_t)sm.gbase[d] + (i - sm.bin_start[d]);
        keys_out[out] = kk;
        vals_out[out] = sm.vals[i];
    }
}

static inline int sort_nblocks(int64_t n) { return (int)((n + SORT_TILE - 1) / SORT_TILE); }

extern "C" uint64_t rs_radix_sort_workspace_bytes(int64_t n) {
    const uint64_t nb = (uint64_t)(sort_nblocks(n) > 0 ? sort_nblocks(n) : 1);
    return (uint64_t)RADIX * nb * sizeof(int32_t) + RADIX * sizeof(int32_t) + 256;
}

template <typename KeyT>
static int radix_sort_impl(int64_t n_bound, const int32_t *n_dev, int begin_bit, int end_bit, KeyT *keys_a,
                           KeyT *keys_b, int32_t *vals_a, int32_t *vals_b, void *workspace, uint64_t workspace_bytes,
                           int32_t *result_in_b, cudaStream_t s) {
    if (result_in_b)
        *result_in_b = 0;
    if (n_bound <= 0 || end_bit <= begin_bit)
        return 0;
    RS_CHECK(keys_a && keys_b && vals_a && vals_b && workspace, "rs_radix_sort_pairs: null pointer");
    RS_CHECK(workspace_bytes >= rs_radix_sort_workspace_bytes(n_bound),
             "rs_radix_sort_pairs: workspace too small (%llu < %llu)", (unsigned long long)workspace_bytes,
             (unsigned long long)rs_radix_sort_workspace_bytes(n_bound));
    RS_CHECK(n_bound < ((int64_t)1 << 31), "rs_radix_sort_pairs: n must fit in int32");
    const int nb = sort_nblocks(n_bound);
    int32_t *hist = reinterpret_cast<int32_t *>(workspace);
    int32_t *bin_tot = hist + (size_t)RADIX * nb;
    static bool attr_set[2] = {false, false};
    const int which = sizeof(KeyT) == 8 ? 1 : 0;
    const size_t smem = sizeof(SortSmem<KeyT>);
    if (!attr_set[which]) {
        RS_CUDA(cudaFuncSetAttribute(sort_scatter_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        attr_set[which] = true;
    }
    KeyT *kin = keys_a, *kout = keys_b;
    int32_t *vin = vals_a, *vout = vals_b;
    int passes = 0;
    for (int shift = begin_bit; shift < end_bit; shift += RADIX_BITS) {
        const int bits = min(RADIX_BITS, end_bit - shift);
        const uint32_t mask = (1u << bits) - 1u;
        sort_hist_kernel<KeyT><<<nb, SORT_THREADS, 0, s>>>(kin, n_bound, n_dev, shift, mask, hist, nb);
        RS_LAUNCH_CHECK("sort_hist_kernel");
        sort_scan_rows_kernel<<<RADIX, SORT_THREADS, 0, s>>>(hist, nb, n_bound, n_dev, bin_tot);
        RS_LAUNCH_CHECK("sort_scan_rows_kernel");
        sort_scatter_kernel<KeyT><<<nb, SORT_THREADS, smem, s>>>(kin, vin, kout, vout, n_bound, n_dev, shift, mask,
                                                                 hist, nb, bin_tot);
        RS_LAUNCH_CHECK("sort_scatter_kernel");
        KeyT *tk = kin;
        kin = kout;
        kout = tk;
        int32_t *tv = vin;
        vin = vout;
        vout = tv;
        ++passes;
    }
    if (result_in_b)
        *result_in_b = passes & 1;
    return 0;
}

extern "C" int rs_radix_sort_pairs(const rs_sort_args *a, rs_stream_t stream) {
    RS_CHECK(a != nullptr, "rs_radix_sort_pairs: null args");
    RS_CHECK(a->begin_bit >= 0 && a->end_bit <= 64 && a->begin_bit <= a->end_bit, "rs_radix_sort_pairs: bad bit range");
    return radix_sort_impl<uint64_t>(a->n, a->n_dev, a->begin_bit, a->end_bit, reinterpret_cast<uint64_t *>(a->keys_a),
                                     reinterpret_cast<uint64_t *>(a->keys_b), a->vals_a, a->vals_b, a->workspace,
                                     a->workspace_bytes, a->result_in_b, (cudaStream_t)stream);
}

int rs_sort_pairs_u32_internal(int64_t n_bound, const int32_t *n_dev, int begin_bit, int end_bit, uint32_t *keys_a,
                               uint32_t *keys_b, int32_t *vals_a, int32_t *vals_b, void *workspace,
                               uint64_t workspace_bytes, int32_t *result_in_b, cudaStream_t s) {
    return radix_sort_impl<uint32_t>(n_bound, n_dev, begin_bit, end_bit, keys_a, keys_b, vals_a, vals_b, workspace,
                                     workspace_bytes, result_in_b, s);
}

// internal entry for the fused frame path
int rs_sort_pairs_u64_internal(int64_t n_bound, const int32_t *n_dev, int begin_bit, int end_bit, uint64_t *keys_a,
                               uint64_t *keys_b, int32_t *vals_a, int32_t *vals_b, void *workspace,
                               uint64_t workspace_bytes, int32_t *result_in_b, cudaStream_t s) {
    return radix_sort_impl<uint64_t>(n_bound, n_dev, begin_bit, end_bit, keys_a, keys_b, vals_a, vals_b, workspace,
                                     workspace_bytes, result_in_b, s);
}
