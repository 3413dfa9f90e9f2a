// Hand-written stable LSD radix sort of (key, int32 value) pairs -- replaces the cub::DeviceRadixSort::SortPairs call
// of csrc/IntersectTile.cu:296-339.  Used twice per frame by the depth-ordered binning of isect.cu (u32 depth keys,
// u32 image|tile keys) and exported as rs_radix_sort_pairs (u64 keys, the reference's image|tile|depth format).
//
// One-sweep organisation, 8-bit digits:
//   sort_hist_kernel   ONE read of the keys builds the 256-bin digit histogram of EVERY pass (LSD passes only permute the
//                      keys, so all global histograms are known up front); it also clears the look-back state.
//   sort_pass_kernel   one launch per digit.  A persistent CTA takes tiles of 4096 pairs in ticket order, ranks them
//                      stably in shared memory (warp match + per-warp digit counters, warp-striped order), obtains the
//                      number of equal-digit pairs in all earlier tiles by DECOUPLED LOOK-BACK (each tile publishes
//                      aggregate / inclusive counts per digit in one 32-bit word; a tile only ever waits on tiles with a
//                      smaller ticket, which are already running), re-orders the tile in shared memory so that equal
//                      digits are contiguous and writes the runs to their final positions.
// Per pass every pair is read once and written once: algorithmic HBM bytes = n * 2 * (sizeof(key) + 4), plus one extra
// key read for the histogram kernel per SORT (not per pass).
// Stability: tile order = ticket order = ascending index; ranks preserve index order inside a tile.
// A device-side pair count (`n_dev`) makes the whole sort launchable without knowing n on the host; grids are sized by
// the SM count, never by the capacity.
#include "common.cuh"

#include "sort_ws.cuh"

#ifndef SORT_LB_MODE
#define SORT_LB_MODE 0
#endif

static_assert(RADIX == SORT_THREADS, "one thread per digit in the block-level scans");

#define LB_AGGREGATE 0x40000000u
#define LB_INCLUSIVE 0x80000000u
#define LB_VALUE 0x3fffffffu
#define LB_WINDOW 8

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

// ---------------------------------------------------------------------------------------------------------------------
// histograms of all passes + look-back reset.  ws[WS_HIST..] and ws[WS_TICKET..] are zeroed by a memset before.
// ---------------------------------------------------------------------------------------------------------------------
// Four privatised copies of every pass histogram (copy = lane & 3): digits that are constant across a warp (the high bits of
// tile ids in emission order, zero upper key bytes) then serialise 8 deep instead of 32 deep, without any per-key
// shuffle / ballot to detect them -- 2 instructions per key and pass (shift-and-mask, ATOMS).
#define SORT_HIST_COPIES 4
template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const KeyT *__restrict__ keys, int64_t n_bound, const int32_t *__restrict__ n_dev, int begin_bit,
                 int end_bit, int passes, int nb_stride, uint32_t *__restrict__ ws) {
    __shared__ unsigned int h[SORT_MAX_PASSES][SORT_HIST_COPIES][RADIX];
    const int64_t n = sort_count(n_bound, n_dev);
    for (int i = threadIdx.x; i < passes * SORT_HIST_COPIES * RADIX; i += SORT_THREADS)
        (&h[0][0][0])[i] = 0u;
    __syncthreads();
    const int copy = threadIdx.x & (SORT_HIST_COPIES - 1);
    const int width = sort_digit_width(end_bit - begin_bit);
    constexpr int VEC = 16 / sizeof(KeyT); // keys per 128-bit load
    const int64_t n_vec = n / VEC;
    const bool aligned = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
    auto count_key = [&](KeyT key) {
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + p * width;
            const int bits = min(width, end_bit - shift);
            atomicAdd(&h[p][copy][(uint32_t)(key >> shift) & ((1u << bits) - 1u)], 1u);
        }
    };
    if (aligned) {
        for (int64_t v = (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * SORT_THREADS) {
            const uint4 q = reinterpret_cast<const uint4 *>(keys)[v];
            if constexpr (sizeof(KeyT) == 4) {
                count_key((KeyT)q.x);
                count_key((KeyT)q.y);
                count_key((KeyT)q.z);
                count_key((KeyT)q.w);
            } else {
                count_key((KeyT)(((unsigned long long)q.y << 32) | q.x));
                count_key((KeyT)(((unsigned long long)q.w << 32) | q.z));
            }
        }
    }
    // tail (or everything, for a misaligned key array)
    for (int64_t i = (aligned ? n_vec * VEC : 0) + (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * SORT_THREADS)
        count_key(keys[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += SORT_THREADS) {
        const int p = i / RADIX, d = i - p * RADIX;
        unsigned int v = 0;
#pragma unroll
        for (int c = 0; c < SORT_HIST_COPIES; ++c)
            v += h[p][c][d];
        if (v)
            atomicAdd(&ws[WS_HIST + p * RADIX + d], v);
    }
    // reset the look-back words of the tiles this sort will use
    const int64_t words = (n + SORT_TILE - 1) / SORT_TILE * RADIX;
    for (int p = 0; p < passes; ++p) {
        uint32_t *lb = ws + WS_LOOKBACK + (size_t)p * nb_stride * RADIX;
        for (int64_t i = (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x; i < words;
             i += (int64_t)gridDim.x * SORT_THREADS)
            lb[i] = 0u;
    }
}

template <typename KeyT> struct SortSmem {
    // the re-ordered tile: 32-bit keys travel with their value as ONE 64-bit word (one STS.64 / LDS.64 per pair instead of
    // two stores and two loads: the pass kernel is bound by the shared-memory / MIO instruction queue, not by bandwidth);
    // 64-bit keys keep two arrays
    KeyT keys[sizeof(KeyT) == 4 ? 1 : SORT_TILE];
    int32_t vals[sizeof(KeyT) == 4 ? 1 : SORT_TILE];
    alignas(16) uint2 kv[sizeof(KeyT) == 4 ? SORT_TILE : 1];
    alignas(16) int32_t vals_stage[SORT_TILE]; // the tile's values in input order, landed by 16-byte cp.async while the keys are ranked
    int wh[SORT_WARPS][RADIX]; // per-warp digit counters -> first slot of (warp, digit) inside the re-ordered tile
    int delta[RADIX];          // global index of slot i of the re-ordered tile = delta[digit] + i
    int real[RADIX];           // digit counts of this tile without padding
    int gstart[RADIX];         // exclusive scan of the global digit histogram
    int wsum[SORT_WARPS];
    int tile;
};

__device__ __forceinline__ void sort_cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void sort_cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// resident CTAs per SM the pass kernel is compiled for (register cap) and launched with: 58 KB (u32) / 74 KB (u64) of
// shared memory per CTA at 16 items per thread; larger tiles (-DSORT_ITEMS=24|32 experiments) leave room for fewer CTAs
#define SORT_CTAS_PER_SM(KeyT) (SORT_ITEMS <= 16 ? (sizeof(KeyT) == 4 ? 3 : 2) : (SORT_ITEMS <= 24 && sizeof(KeyT) == 4 ? 2 : 1))
template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS, SORT_CTAS_PER_SM(KeyT))
sort_pass_kernel(const KeyT *__restrict__ keys_in, const int32_t *__restrict__ vals_in, KeyT *__restrict__ keys_out,
                 int32_t *__restrict__ vals_out, int64_t n_bound, const int32_t *__restrict__ n_dev, int shift,
                 uint32_t mask, int pass, int nb_stride, uint32_t *__restrict__ ws) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem<KeyT> &sm = *reinterpret_cast<SortSmem<KeyT> *>(sort_smem_raw);
    const int64_t n = sort_count(n_bound, n_dev);
    const int nb_eff = (int)((n + SORT_TILE - 1) / SORT_TILE);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = rs_lanemask_lt();
    uint32_t *lookback = ws + WS_LOOKBACK + (size_t)pass * nb_stride * RADIX;

    // exclusive scan of this pass's global histogram: where each digit's output range starts
    {
        const int g = (int)ws[WS_HIST + pass * RADIX + threadIdx.x];
        const int ex = block_excl_scan_256(g, sm.wsum, nullptr);
        sm.gstart[threadIdx.x] = ex;
    }

    while (true) {
        __syncthreads(); // previous tile fully written out; sm.tile free
        if (threadIdx.x == 0)
            sm.tile = (int)atomicAdd(&ws[WS_TICKET + pass], 1u);
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            sm.wh[w][threadIdx.x] = 0;
        __syncthreads();
        const int tile = sm.tile;
        if (tile >= nb_eff)
            break;
        const int64_t tile_start = (int64_t)tile * SORT_TILE;
        const int count = (int)min((int64_t)SORT_TILE, n - tile_start);

        // warp-striped load: warp w owns [w*512, (w+1)*512), item k of lane l sits at k*32 + l
        KeyT key[SORT_ITEMS];
        const int wbase = warp * (32 * SORT_ITEMS);
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const int local = wbase + k * 32 + lane;
            key[k] = (local < count) ? keys_in[tile_start + local]
                                     : ~(KeyT)0; // sorts to the very end of the tile, never written out
        }

        // values: asynchronous copy into shared memory (no registers held across the ranking)
        if (vals_in != nullptr) {
            const int32_t *src = vals_in + tile_start;
            if (count == SORT_TILE && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
                for (int q = 0; q < SORT_TILE / (SORT_THREADS * 4); ++q) {
                    const int e = (q * SORT_THREADS + threadIdx.x) * 4;
                    sort_cp_async16(&sm.vals_stage[e], src + e);
                }
            } else {
                for (int e = threadIdx.x; e < count; e += SORT_THREADS)
                    sort_cp_async4(&sm.vals_stage[e], src + e);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }

        // Stable ranks.  All SORT_ITEMS match operations are independent (pipelined); the per-warp digit counters are
        // bumped with shared-memory atomics by the leader lane of each peer group: atomics of one warp reach the LSU in
        // program order, so round k is counted before round k+1 without any read-modify-write dependency chain.
        unsigned peers[SORT_ITEMS];
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k)
            peers[k] = __match_any_sync(0xffffffffu, (uint32_t)(key[k] >> shift) & mask);
        int rank[SORT_ITEMS];
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & mask;
            const int leader = __ffs(peers[k]) - 1;
            int old = 0;
            if (lane == leader)
                old = atomicAdd(&sm.wh[warp][d], __popc(peers[k]));
            __syncwarp();
            rank[k] = old;
        }
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const int leader = __ffs(peers[k]) - 1;
            rank[k] = __shfl_sync(0xffffffffu, rank[k], leader) + __popc(peers[k] & lt);
        }
        __syncthreads();

        // thread t handles digit t: per-warp exclusive offsets, tile-level digit start, published aggregate
        int bin_start = 0;
        {
            int sum = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) {
                const int c = sm.wh[w][threadIdx.x];
                sm.wh[w][threadIdx.x] = sum;
                sum += c;
            }
            // padding keys (all ones) of a ragged last tile were counted in the top digit: take them out of the counts
            // that are published (they are the last entries of that digit, so local ranks are unaffected)
            int real = sum;
            if (count < SORT_TILE && threadIdx.x == (int)(((uint32_t)((~(KeyT)0) >> shift)) & mask))
                real = sum - (SORT_TILE - count);
            st_volatile_u32(lookback + (size_t)tile * RADIX + threadIdx.x,
                            (tile == 0 ? LB_INCLUSIVE : LB_AGGREGATE) | (uint32_t)real);
            sm.real[threadIdx.x] = real;
            bin_start = block_excl_scan_256(sum, sm.wsum, nullptr); // (contains the barrier that publishes sm.real)
            // first slot of every (warp, digit) group inside the re-ordered tile: one load per pair in the scatter below
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w)
                sm.wh[w][threadIdx.x] += bin_start;
        }
        // Decoupled look-back, one digit per thread.  Windowed: LB_WINDOW independent loads per L2 round trip -- when all
        // tiles of a wave start together, the inclusive prefix can only advance one window per round trip, and that chain
        // of round trips (not bandwidth, not instruction issue) is what bounds a pass.
        {
            const int d = threadIdx.x;
            int excl = 0;
#if SORT_LB_MODE == 1 // timing experiment only: no waiting at all (results are wrong)
            if (false) {
#elif SORT_LB_MODE == 2 // every predecessor's aggregate is summed directly: no chain of inclusive prefixes to wait for
            if (tile > 0) {
                for (int t0 = tile - 1; t0 >= 0; t0 -= 16) {
                    uint32_t v[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj)
                        v[jj] = (t0 - jj >= 0) ? ld_volatile_u32(lookback + (size_t)(t0 - jj) * RADIX + d) : LB_AGGREGATE;
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        while ((v[jj] & (LB_AGGREGATE | LB_INCLUSIVE)) == 0u)
                            v[jj] = ld_volatile_u32(lookback + (size_t)(t0 - jj) * RADIX + d);
                    }
                    bool stop = false;
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        if (!stop) {
                            excl += (int)(v[jj] & LB_VALUE);
                            stop = (v[jj] & LB_INCLUSIVE) != 0u;
                        }
                    }
                    if (stop)
                        break;
                }
                st_volatile_u32(lookback + (size_t)tile * RADIX + d, LB_INCLUSIVE | (uint32_t)(excl + sm.real[d]));
            }
            if (false) {
#else
            if (tile > 0) {
#endif
                int t = tile - 1;
                bool found = false;
                while (!found) {
                    uint32_t v[LB_WINDOW];
#pragma unroll
                    for (int jj = 0; jj < LB_WINDOW; ++jj)
                        v[jj] = (t - jj >= 0) ? ld_volatile_u32(lookback + (size_t)(t - jj) * RADIX + d)
                                              : LB_INCLUSIVE; // before tile 0: inclusive prefix 0
                    int used = 0;
#pragma unroll
                    for (int jj = 0; jj < LB_WINDOW; ++jj) {
                        if (!found && used == jj && (v[jj] & (LB_AGGREGATE | LB_INCLUSIVE)) != 0u) {
                            excl += (int)(v[jj] & LB_VALUE);
                            ++used;
                            found = (v[jj] & LB_INCLUSIVE) != 0u;
                        }
                    }
                    t -= used;
                    if (used == 0)
                        __nanosleep(200); // predecessor not published yet: leave the issue slots to other warps / frames
                }
                st_volatile_u32(lookback + (size_t)tile * RADIX + d, LB_INCLUSIVE | (uint32_t)(excl + sm.real[d]));
            }
            // global index of slot i of the re-ordered tile = delta[digit] + i
            sm.delta[d] = sm.gstart[d] + excl - bin_start;
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory"); // this thread's value copies have landed ...
        __syncthreads();                                       // ... and so have everyone else's

#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & mask;
            const int pos = sm.wh[warp][d] + rank[k];
            const int local = wbase + k * 32 + lane;
            const int32_t v = vals_in != nullptr ? sm.vals_stage[min(local, SORT_TILE - 1)] : (int32_t)(tile_start + local);
            if constexpr (sizeof(KeyT) == 4) {
                sm.kv[pos] = make_uint2((uint32_t)key[k], (uint32_t)v); // padding keys land behind slot `count`: never read
            } else {
                sm.keys[pos] = key[k];
                if (local < count)
                    sm.vals[pos] = v;
            }
        }
        __syncthreads();

        for (int i = threadIdx.x; i < count; i += SORT_THREADS) {
            KeyT kk;
            int32_t vv;
            if constexpr (sizeof(KeyT) == 4) {
                const uint2 p = sm.kv[i];
                kk = (KeyT)p.x;
                vv = (int32_t)p.y;
            } else {
                kk = sm.keys[i];
                vv = sm.vals[i];
            }
            const uint32_t d = (uint32_t)(kk >> shift) & mask;
            const int64_t out = (int64_t)(sm.delta[d] + i);
            keys_out[out] = kk;
            vals_out[out] = vv;
        }
    }
}

static inline int sort_nblocks(int64_t n) { return (int)((n + SORT_TILE - 1) / SORT_TILE); }

extern "C" uint64_t rs_radix_sort_workspace_bytes(int64_t n) {
    const uint64_t nb = (uint64_t)(sort_nblocks(n) > 0 ? sort_nblocks(n) : 1);
    return ((uint64_t)WS_LOOKBACK + (uint64_t)SORT_MAX_PASSES * nb * RADIX) * sizeof(uint32_t) + 256;
}

// keys_in/vals_in are only read (vals_in == nullptr: values are the indices 0..n-1); pass p writes buffer p & 1, so the
// result ends in buffer (passes - 1) & 1.  buf 0 must not alias the input; buf 1 may.
template <typename KeyT>
static int radix_sort_impl(int64_t n_bound, const int32_t *n_dev, int begin_bit, int end_bit, const KeyT *keys_in,
                           const int32_t *vals_in, KeyT *kbuf0, int32_t *vbuf0, KeyT *kbuf1, int32_t *vbuf1,
                           void *workspace, uint64_t workspace_bytes, int *passes_out, cudaStream_t s,
                           bool hist_ready = false) {
    if (passes_out)
        *passes_out = 0;
    if (n_bound <= 0 || end_bit <= begin_bit)
        return 0;
    const int passes = sort_num_passes(end_bit - begin_bit);
    const int width = sort_digit_width(end_bit - begin_bit);
    RS_CHECK(passes <= SORT_MAX_PASSES, "rs_radix_sort_pairs: too many key bits");
    RS_CHECK(keys_in && kbuf0 && vbuf0 && workspace && (passes < 2 || (kbuf1 && vbuf1)),
             "rs_radix_sort_pairs: null pointer");
    RS_CHECK(workspace_bytes >= rs_radix_sort_workspace_bytes(n_bound),
             "rs_radix_sort_pairs: workspace too small (%llu < %llu)", (unsigned long long)workspace_bytes,
             (unsigned long long)rs_radix_sort_workspace_bytes(n_bound));
    RS_CHECK(n_bound < (int64_t)LB_VALUE, "rs_radix_sort_pairs: n must be below 2^30");
    const int nb = sort_nblocks(n_bound);
    uint32_t *ws = reinterpret_cast<uint32_t *>(workspace);
    static RsPerDevice attr_set[2];
    const int which = sizeof(KeyT) == 8 ? 1 : 0;
    const size_t smem = sizeof(SortSmem<KeyT>);
    if (!rs_dev_done(attr_set[which])) {
        RS_CUDA(cudaFuncSetAttribute(sort_pass_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rs_dev_mark(attr_set[which]);
    }
    const int sms = rs_num_sms();
    if (!hist_ready) { // else: the producer of the keys already filled ws (rs_sort_ws_prepare + its own histogramming)
        RS_CUDA(cudaMemsetAsync(ws, 0, (size_t)WS_LOOKBACK * sizeof(uint32_t), s));
        const int hist_grid = (int)min((int64_t)sms * 6, (n_bound + SORT_THREADS * 4 - 1) / (SORT_THREADS * 4));
        sort_hist_kernel<KeyT><<<hist_grid, SORT_THREADS, 0, s>>>(keys_in, n_bound, n_dev, begin_bit, end_bit, passes, nb,
                                                                  ws);
        RS_LAUNCH_CHECK("sort_hist_kernel");
    }
    const int pass_grid = min(nb, sms * SORT_CTAS_PER_SM(KeyT));
    const KeyT *kin = keys_in;
    const int32_t *vin = vals_in;
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * width;
        const int bits = min(width, end_bit - shift);
        const uint32_t mask = (1u << bits) - 1u;
        KeyT *kout = (p & 1) ? kbuf1 : kbuf0;
        int32_t *vout = (p & 1) ? vbuf1 : vbuf0;
        sort_pass_kernel<KeyT><<<pass_grid, SORT_THREADS, smem, s>>>(kin, vin, kout, vout, n_bound, n_dev, shift, mask, p,
                                                                     nb, ws);
        RS_LAUNCH_CHECK("sort_pass_kernel");
        kin = kout;
        vin = vout;
    }
    if (passes_out)
        *passes_out = passes;
    return 0;
}

extern "C" int rs_radix_sort_pairs(const rs_sort_args *a, rs_stream_t stream) {
    RS_CHECK(a != nullptr, "rs_radix_sort_pairs: null args");
    RS_CHECK(a->begin_bit >= 0 && a->end_bit <= 64 && a->begin_bit <= a->end_bit, "rs_radix_sort_pairs: bad bit range");
    if (a->result_in_b)
        *a->result_in_b = 0;
    RS_CHECK(a->n <= 0 || a->end_bit == a->begin_bit || a->vals_a != nullptr, "rs_radix_sort_pairs: null pointer");
    int passes = 0;
    // double-buffer semantics of the reference's cub call: input in a, pass 0 writes b, pass 1 writes a, ...
    const int e = radix_sort_impl<uint64_t>(a->n, a->n_dev, a->begin_bit, a->end_bit,
                                            reinterpret_cast<const uint64_t *>(a->keys_a), a->vals_a,
                                            reinterpret_cast<uint64_t *>(a->keys_b), a->vals_b,
                                            reinterpret_cast<uint64_t *>(a->keys_a), a->vals_a, a->workspace,
                                            a->workspace_bytes, &passes, (cudaStream_t)stream);
    if (e == 0 && a->result_in_b)
        *a->result_in_b = passes & 1;
    return e;
}

// same as rs_radix_sort_pairs for 32-bit keys (keys_a / keys_b point to uint32 arrays): the sort the binning path runs
extern "C" int rs_radix_sort_pairs32(const rs_sort_args *a, rs_stream_t stream) {
    RS_CHECK(a != nullptr, "rs_radix_sort_pairs32: null args");
    RS_CHECK(a->begin_bit >= 0 && a->end_bit <= 32 && a->begin_bit <= a->end_bit, "rs_radix_sort_pairs32: bad bit range");
    if (a->result_in_b)
        *a->result_in_b = 0;
    RS_CHECK(a->n <= 0 || a->end_bit == a->begin_bit || a->vals_a != nullptr, "rs_radix_sort_pairs32: null pointer");
    int passes = 0;
    const int e = radix_sort_impl<uint32_t>(a->n, a->n_dev, a->begin_bit, a->end_bit,
                                            reinterpret_cast<const uint32_t *>(a->keys_a), a->vals_a,
                                            reinterpret_cast<uint32_t *>(a->keys_b), a->vals_b,
                                            reinterpret_cast<uint32_t *>(a->keys_a), a->vals_a, a->workspace,
                                            a->workspace_bytes, &passes, (cudaStream_t)stream);
    if (e == 0 && a->result_in_b)
        *a->result_in_b = passes & 1;
    return e;
}

// internal entry for the binning path (isect.cu): u32 keys, read-only input, optional implicit index values
int rs_sort_pairs_u32_internal(int64_t n_bound, const int32_t *n_dev, int begin_bit, int end_bit, const uint32_t *keys_in,
                               const int32_t *vals_in, uint32_t *kbuf0, int32_t *vbuf0, uint32_t *kbuf1, int32_t *vbuf1,
                               void *workspace, uint64_t workspace_bytes, int *passes, cudaStream_t s, bool hist_ready) {
    return radix_sort_impl<uint32_t>(n_bound, n_dev, begin_bit, end_bit, keys_in, vals_in, kbuf0, vbuf0, kbuf1, vbuf1,
                                     workspace, workspace_bytes, passes, s, hist_ready);
}

// zero the histogram / ticket header of a sort workspace: to be enqueued before a kernel that accumulates the digit
// histograms itself (hist_ready = true above); that kernel must also clear the look-back words it will need.
int rs_sort_ws_prepare(void *workspace, cudaStream_t s) {
    RS_CUDA(cudaMemsetAsync(workspace, 0, (size_t)WS_LOOKBACK * sizeof(uint32_t), s));
    return 0;
}
