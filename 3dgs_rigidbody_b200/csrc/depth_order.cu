// Depth order of the visible (image, Gaussian) elements: the first half of the depth-ordered binning of isect.cu.
//
// What is needed: the visible elements listed in ascending (depth bits, flatten index) order -- the order in which a
// stable sort of the reference's (image | tile | depth) keys breaks ties inside a tile (csrc/IntersectTile.cu:95-113,
// 296-339).  Round 1 produced it with four one-sweep LSD passes over all E elements; at E = 1 M every pass is a single
// wave of CTAs chained through a look-back and costs ~23 us whatever the bandwidth (profiles/r01_sort_microbench.txt).
//
// Here the order is produced by a bucket sort whose steps have no serial chain at all:
//   K0  min / max of the depth bits of the visible elements, and their number V                     (E reads)
//   K1  bucket histogram: bucket = (bits - min) >> shift, 2^16 .. 2^20 buckets (~32 elements per bucket beyond 2 M elements)
//       over the ACTUAL depth range of this frame, so the resolution adapts to the scene              (E reads, V atomics)
//   K2  exclusive scan of the bucket counts by one CTA; buckets are grouped into sort groups of ~DORD_TARGET elements
//       (group g starts at the first bucket boundary at or after g * DORD_TARGET)
//   K3  scatter: every visible element takes the next free slot of its bucket (atomic cursor) and stores the 64-bit
//       composite (depth bits << 32 | flatten index)                                                  (V atomics, 8 B writes)
//   K4  one CTA per sort group: the group's composites are ordered in shared memory (they are unique, so ascending composite
//       order IS (depth, index) order, whatever order the atomics of K3 produced) by a second-level counting sort over 2048
//       sub-buckets of the group's own key range plus an insertion sort of the few multiply-occupied sub-buckets; groups with
//       long runs of near-identical depths take a bitonic network instead.  A group that does not fit in shared memory (thousands of elements with depth bits inside one
//       bucket, e.g. a wall facing an orthographic camera) is sorted by the same CTA in global memory with a stable LSD
//       radix sort over the bits that actually vary -- slow, but correct for any input.
// Output: elems[0 .. V) and V (device side).  Everything is sized by the SM count / the element bound; no host sync.
#include "common.cuh"

// Bucket count: 65536 up to 2 M elements, then ~32 elements per bucket up to 2^20 buckets.  Too few buckets and the atomics
// of K1 / K3 pile up on too few L2 addresses (8192 buckets at 1 M elements: count 14 -> 37 us, scatter 19 -> 40 us,
// profiles/r02_depth_order_experiments.txt); the look-back scan does not care about the bucket count.
#define DORD_BUCKET_BITS_MIN 16
#define DORD_BUCKET_BITS_MAX 20
#define DORD_BUCKETS_MAX (1 << DORD_BUCKET_BITS_MAX)
static inline int dord_bucket_bits(int64_t n_elems) {
    int bits = DORD_BUCKET_BITS_MIN;
    while (bits < DORD_BUCKET_BITS_MAX && ((int64_t)32 << bits) < n_elems)
        ++bits;
    return bits;
}
#define DORD_TARGET 896   // elements per sort group (plus the tail of the bucket that crosses the boundary)
#define DORD_CAP 4096     // composites a CTA sorts in shared memory
#define DORD_THREADS 256

struct DordHeader {          // zero-initialised by one memset per call
    unsigned int inv_min;    // max over visible elements of ~bits  (min = ~inv_min)
    unsigned int max;        // max of bits
    unsigned int n_visible;  // V
    unsigned int shift;      // bucket = (bits - min) >> shift          (written by K2's prologue ... see dord_shift)
    unsigned int n_groups;
    unsigned int _pad[3];
};

__device__ __forceinline__ unsigned int dord_shift(unsigned int kmin, unsigned int kmax, int bucket_bits) {
    const unsigned int range = kmax - kmin; // buckets must cover [0, range]
    const int width = 32 - __clz(range | 1u);
    return (unsigned int)max(0, width - bucket_bits);
}

// ---- K0 -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DORD_THREADS)
rs_dord_minmax_kernel(int64_t n_elems, const float *__restrict__ depths, const int32_t *__restrict__ tiles,
                      DordHeader *__restrict__ hdr) {
    unsigned int inv_min = 0u, mx = 0u, cnt = 0u;
    for (int64_t i = (int64_t)blockIdx.x * DORD_THREADS + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * DORD_THREADS) {
        if (tiles[i] > 0) {
            const unsigned int k = __float_as_uint(depths[i]);
            inv_min = max(inv_min, ~k);
            mx = max(mx, k);
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inv_min = max(inv_min, __shfl_xor_sync(0xffffffffu, inv_min, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    // one atomic triple per CTA: same-address atomics serialise in L2, so per-warp atomics would dominate the kernel
    __shared__ unsigned int s_red[3][DORD_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_red[0][warp] = inv_min;
        s_red[1][warp] = mx;
        s_red[2][warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < DORD_THREADS / 32; ++w) {
            inv_min = max(inv_min, s_red[0][w]);
            mx = max(mx, s_red[1][w]);
            cnt += s_red[2][w];
        }
        if (cnt != 0u) {
            atomicMax(&hdr->inv_min, inv_min);
            atomicMax(&hdr->max, mx);
            atomicAdd(&hdr->n_visible, cnt);
        }
    }
}

// ---- K1 -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DORD_THREADS)
rs_dord_count_kernel(int64_t n_elems, const float *__restrict__ depths, const int32_t *__restrict__ tiles,
                     const DordHeader *__restrict__ hdr, unsigned int *__restrict__ counts, int bucket_bits) {
    const unsigned int kmin = ~hdr->inv_min, shift = dord_shift(kmin, hdr->max, bucket_bits);
    for (int64_t i = (int64_t)blockIdx.x * DORD_THREADS + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * DORD_THREADS) {
        if (tiles[i] > 0)
            atomicAdd(&counts[(__float_as_uint(depths[i]) - kmin) >> shift], 1u);
    }
}

// ---- K2: counts -> bucket cursors (exclusive prefix, in place) + group starts --------------------------------------------
// 64 CTAs x 1024 buckets (4 consecutive buckets per thread, one 128-bit load / store each).  The CTA totals are chained by a
// decoupled look-back over one status word per CTA ({flag : 2 | value : 30}: 1 = the CTA's own total, 2 = inclusive prefix;
// one warp inspects up to 32 predecessors per round trip, so the chain is two steps long).  A single-CTA scan of the same
// 256 KB took 20 us (one SM's bandwidth and a 64-step dependent chain); this takes the latency of one wave.
#define DORD_SCAN_THREADS 256
#define DORD_SCAN_PER_CTA (DORD_SCAN_THREADS * 4)
#define DORD_SCAN_AGG 0x40000000u
#define DORD_SCAN_PREFIX 0x80000000u
#define DORD_SCAN_VALUE 0x3fffffffu
__global__ void __launch_bounds__(DORD_SCAN_THREADS)
rs_dord_scan_kernel(DordHeader *__restrict__ hdr, unsigned int *__restrict__ counts, unsigned int *__restrict__ state,
                    unsigned int *__restrict__ group_start, int32_t *__restrict__ n_sorted_out, int bucket_bits) {
    __shared__ unsigned int warp_tot[DORD_SCAN_THREADS / 32];
    __shared__ unsigned int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int f0 = blockIdx.x * DORD_SCAN_PER_CTA + threadIdx.x * 4;
    const uint4 c = *reinterpret_cast<const uint4 *>(counts + f0);
    const unsigned int sum = c.x + c.y + c.z + c.w;
    unsigned int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += n;
    }
    if (lane == 31)
        warp_tot[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < DORD_SCAN_THREADS / 32; ++w) {
        const unsigned int t = warp_tot[w];
        wbase += (w < warp) ? t : 0u;
        total += t;
    }
    if (warp == 0) {
        volatile unsigned int *st = state;
        unsigned int prefix = 0;
        if (blockIdx.x > 0) {
            if (lane == 0)
                st[blockIdx.x] = DORD_SCAN_AGG | total;
            int look = (int)blockIdx.x - 1;
            while (true) {
                const int j = look - lane;
                unsigned int v;
                do {
                    v = j >= 0 ? st[j] : DORD_SCAN_PREFIX;
                } while (__any_sync(0xffffffffu, (v & (DORD_SCAN_AGG | DORD_SCAN_PREFIX)) == 0u));
                const unsigned int pm = __ballot_sync(0xffffffffu, (v & DORD_SCAN_PREFIX) != 0u);
                unsigned int val = v & DORD_SCAN_VALUE;
                if (pm != 0u && lane > (__ffs(pm) - 1))
                    val = 0u; // behind the nearest published inclusive prefix
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
            __threadfence();
            st[blockIdx.x] = DORD_SCAN_PREFIX | (prefix + total);
            s_base = prefix;
        }
    }
    __syncthreads();
    const unsigned int V = hdr->n_visible;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        group_start[0] = 0u;
        hdr->n_groups = V / DORD_TARGET + 1u;
        group_start[V / DORD_TARGET + 1u] = V;
        hdr->shift = dord_shift(~hdr->inv_min, hdr->max, bucket_bits);
        if (n_sorted_out != nullptr)
            *n_sorted_out = (int32_t)V;
    }
    unsigned int run = s_base + wbase + incl - sum; // elements in all buckets before this thread's first one
    const unsigned int cc[4] = {c.x, c.y, c.z, c.w};
    unsigned int pre[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned int p = run, e = run + cc[j];
        pre[j] = p; // the bucket's cursor for K3
        if (cc[j] != 0u) {
            // every multiple g * TARGET inside (p, e]: group g starts at this bucket's END (first boundary at or after it)
            for (unsigned int g = p / DORD_TARGET + 1u; g * DORD_TARGET <= e; ++g)
                group_start[g] = e;
        }
        run = e;
    }
    *reinterpret_cast<uint4 *>(counts + f0) = make_uint4(pre[0], pre[1], pre[2], pre[3]);
}

// ---- K3 -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DORD_THREADS)
rs_dord_scatter_kernel(int64_t n_elems, const float *__restrict__ depths, const int32_t *__restrict__ tiles,
                       const DordHeader *__restrict__ hdr, unsigned int *__restrict__ cursors,
                       unsigned long long *__restrict__ comp) {
    const unsigned int kmin = ~hdr->inv_min, shift = hdr->shift;
    for (int64_t i = (int64_t)blockIdx.x * DORD_THREADS + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * DORD_THREADS) {
        if (tiles[i] > 0) {
            const unsigned int k = __float_as_uint(depths[i]);
            const unsigned int pos = atomicAdd(&cursors[(k - kmin) >> shift], 1u);
            comp[pos] = ((unsigned long long)k << 32) | (unsigned long long)(unsigned int)i;
        }
    }
}

// ---- K4 -----------------------------------------------------------------------------------------------------------------
// stable LSD radix sort of `n` composites by ONE warp, 8-bit digits over the bytes whose bits vary inside the group;
// src / dst ping-pong in global memory, the result is left in `a` (copied back if needed).  Pathological inputs only.
__device__ void dord_sort_big(unsigned long long *a, unsigned long long *b, unsigned int n, unsigned int *hist /*smem 256*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned long long vary_s;
    if (threadIdx.x == 0)
        vary_s = 0ull;
    __syncthreads();
    // which bits differ from the first composite
    {
        const unsigned long long first = a[0];
        unsigned long long v = 0ull;
        for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
            v |= a[i] ^ first;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v != 0ull)
            atomicOr(&vary_s, v);
    }
    __syncthreads();
    const unsigned long long vary = vary_s;
    unsigned long long *src = a, *dst = b;
    for (int byte = 0; byte < 8; ++byte) {
        if (((vary >> (8 * byte)) & 0xffull) == 0ull)
            continue;
        const int sh = 8 * byte;
        hist[threadIdx.x] = 0u;
        __syncthreads();
        for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
            atomicAdd(&hist[(unsigned int)(src[i] >> sh) & 255u], 1u);
        __syncthreads();
        if (warp == 0) { // exclusive scan of the 256 counters by one warp (8 per lane)
            unsigned int v[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = hist[lane * 8 + j];
                s += v[j];
            }
            unsigned int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += t;
            }
            unsigned int run = incl - s;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hist[lane * 8 + j] = run;
                run += v[j];
            }
            __syncwarp();
            // in-order scatter, 32 composites at a time: equal digits keep their order (rank = earlier lanes with my digit)
            for (unsigned int i0 = 0; i0 < n; i0 += 32) {
                const unsigned int i = i0 + lane;
                const bool live = i < n;
                const unsigned long long c = live ? src[i] : 0ull;
                const unsigned int d = live ? ((unsigned int)(c >> sh) & 255u) : 256u + lane; // dead lanes: unique digits
                const unsigned int peers = __match_any_sync(0xffffffffu, d);
                const unsigned int rank = __popc(peers & rs_lanemask_lt());
                unsigned int base = 0;
                if (live && rank == 0)
                    base = atomicAdd(&hist[d], (unsigned int)__popc(peers));
                base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
                if (live)
                    dst[base + rank] = c;
                __syncwarp();
            }
        }
        __syncthreads();
        unsigned long long *t = src;
        src = dst;
        dst = t;
    }
    if (src != a) {
        for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
            a[i] = src[i];
    }
    __syncthreads();
}

// Fast path of K4 (n <= DORD_FAST): second-level COUNTING sort in shared memory.  The keys of a group span a narrow range
// (its buckets are consecutive), so (key - group min) >> s2 spreads them over DORD_SUBS sub-buckets with < 1 element each
// on average: count, scan, scatter, then order the few sub-buckets that hold more than one composite with an insertion sort
// by the thread that finds the run.  ~80 instructions per element instead of the ~550 of a bitonic network.
#define DORD_FAST 2048
#define DORD_SUBS 2048
#define DORD_RUN_MAX 48 // longer runs of one sub-bucket (near-identical depths): the group takes the bitonic network instead

__device__ __forceinline__ void dord_bitonic(unsigned long long *s, unsigned int n) {
    unsigned int m = 32;
    while (m < n)
        m <<= 1;
    for (unsigned int i = n + threadIdx.x; i < m; i += DORD_THREADS)
        s[i] = ~0ull;
    __syncthreads();
    for (unsigned int k = 2; k <= m; k <<= 1) {
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            for (unsigned int t = threadIdx.x; t < (m >> 1); t += DORD_THREADS) {
                const unsigned int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)); // index with bit j clear
                const unsigned int p = i | j;
                const unsigned long long x = s[i], y = s[p];
                const bool up = (i & k) == 0;
                if ((x > y) == up) {
                    s[i] = y;
                    s[p] = x;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(DORD_THREADS)
rs_dord_sort_kernel(const DordHeader *__restrict__ hdr, const unsigned int *__restrict__ group_start,
                    unsigned long long *__restrict__ comp, unsigned long long *__restrict__ comp_alt,
                    int32_t *__restrict__ elems) {
    __shared__ unsigned long long s[DORD_CAP]; // fast path: [0, FAST) input, [FAST, 2 FAST) output; bitonic: all of it
    __shared__ unsigned int cnt[DORD_SUBS];
    __shared__ unsigned int red[2][DORD_THREADS / 32];
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_flag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int n_groups = hdr->n_groups;
    for (unsigned int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const unsigned int lo = group_start[g], hi = group_start[g + 1];
        if (hi <= lo)
            continue;
        const unsigned int n = hi - lo;
        __syncthreads(); // the previous group of this CTA is fully written out
        if (n > DORD_CAP) {
            dord_sort_big(comp + lo, comp_alt + lo, n, hist);
            for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
                elems[lo + i] = (int32_t)(unsigned int)comp[lo + i];
            continue;
        }
        // load + key range of the group
        unsigned int kmin = 0xffffffffu, kmax = 0u;
        for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS) {
            const unsigned long long c = comp[lo + i];
            s[i] = c;
            const unsigned int k = (unsigned int)(c >> 32);
            kmin = min(kmin, k);
            kmax = max(kmax, k);
        }
        bool fast = n <= DORD_FAST;
        if (fast) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
                kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
            }
            if (lane == 0) {
                red[0][warp] = kmin;
                red[1][warp] = kmax;
            }
            for (unsigned int i = threadIdx.x; i < DORD_SUBS; i += DORD_THREADS)
                cnt[i] = 0u;
            if (threadIdx.x == 0)
                s_flag = 0u;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < DORD_THREADS / 32; ++w) {
                kmin = min(kmin, red[0][w]);
                kmax = max(kmax, red[1][w]);
            }
            const int width = 32 - __clz((kmax - kmin) | 1u);
            const unsigned int s2 = (unsigned int)max(0, width - 11); // (kmax - kmin) >> s2 < DORD_SUBS
            for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
                atomicAdd(&cnt[((unsigned int)(s[i] >> 32) - kmin) >> s2], 1u);
            __syncthreads();
            // exclusive scan of the sub-bucket counts: warp w owns 256 consecutive counters, 32 at a time
            {
                unsigned int c[DORD_SUBS / DORD_THREADS], sum = 0, big = 0;
#pragma unroll
                for (int j = 0; j < DORD_SUBS / DORD_THREADS; ++j) {
                    c[j] = cnt[warp * (DORD_SUBS / 8) + j * 32 + lane];
                    sum += c[j];
                    big = max(big, c[j]);
                }
                if (big > DORD_RUN_MAX)
                    s_flag = 1u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0)
                    red[0][warp] = sum;
                __syncthreads();
                unsigned int carry = 0;
#pragma unroll
                for (int w = 0; w < DORD_THREADS / 32; ++w)
                    carry += (w < warp) ? red[0][w] : 0u;
#pragma unroll
                for (int j = 0; j < DORD_SUBS / DORD_THREADS; ++j) {
                    unsigned int incl = c[j];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o)
                            incl += t;
                    }
                    cnt[warp * (DORD_SUBS / 8) + j * 32 + lane] = carry + incl - c[j];
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            __syncthreads();
            fast = s_flag == 0u;
            if (fast) {
                unsigned long long *out = s + DORD_FAST;
                for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS) {
                    const unsigned long long c = s[i];
                    out[atomicAdd(&cnt[((unsigned int)(c >> 32) - kmin) >> s2], 1u)] = c;
                }
                __syncthreads();
                // order inside the sub-buckets: the thread at the first slot of a run sorts it (runs are disjoint)
                for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS) {
                    const unsigned int sub = ((unsigned int)(out[i] >> 32) - kmin) >> s2;
                    if (i > 0 && (((unsigned int)(out[i - 1] >> 32) - kmin) >> s2) == sub)
                        continue;
                    unsigned int e = i + 1;
                    while (e < n && (((unsigned int)(out[e] >> 32) - kmin) >> s2) == sub)
                        ++e;
                    for (unsigned int a = i + 1; a < e; ++a) { // insertion sort of out[i .. e)
                        const unsigned long long v = out[a];
                        unsigned int q = a;
                        while (q > i && out[q - 1] > v) {
                            out[q] = out[q - 1];
                            --q;
                        }
                        out[q] = v;
                    }
                }
                __syncthreads();
                for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
                    elems[lo + i] = (int32_t)(unsigned int)out[i];
                continue;
            }
        }
        __syncthreads();
        dord_bitonic(s, n);
        for (unsigned int i = threadIdx.x; i < n; i += DORD_THREADS)
            elems[lo + i] = (int32_t)(unsigned int)s[i];
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
namespace {
struct DordLayout {
    size_t hdr, state, counts, group_start, comp, comp_alt, total;
};
inline size_t dord_align(size_t x) { return (x + 255) & ~(size_t)255; }
DordLayout dord_layout(int64_t n_elems) {
    DordLayout L;
    const size_t E = (size_t)(n_elems > 0 ? n_elems : 1);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = dord_align(o + bytes);
        return at;
    };
    L.hdr = take(sizeof(DordHeader));
    L.state = take((size_t)(DORD_BUCKETS_MAX / 1024) * 4); // look-back words of the scan (cleared with the header)
    L.counts = take((size_t)DORD_BUCKETS_MAX * 4);
    L.group_start = take((E / DORD_TARGET + 3) * 4);
    L.comp = take(E * 8);
    L.comp_alt = take(E * 8);
    L.total = o;
    return L;
}
} // namespace

uint64_t rs_depth_order_workspace_bytes(int64_t n_elems) { return dord_layout(n_elems).total; }

// elems_out[0 .. V) = visible elements (tiles[e] > 0) in ascending (depth bits, index) order; *n_sorted_dev = V.
// clears header, look-back words and bucket counters of a depth-order workspace (one memset)
int rs_depth_order_prepare(void *workspace, int64_t n_elems, cudaStream_t s) {
    const DordLayout L = dord_layout(n_elems);
    char *w = reinterpret_cast<char *>(workspace);
    RS_CUDA(cudaMemsetAsync(w + L.hdr, 0, L.counts + ((size_t)4 << dord_bucket_bits(n_elems)) - L.hdr, s));
    return 0;
}
// the three words {max(~bits), max(bits), visible rows} a fused producer accumulates (DordHeader's first fields)
uint32_t *rs_depth_order_stats_ptr(void *workspace, int64_t n_elems) {
    return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(workspace) + dord_layout(n_elems).hdr);
}

int rs_depth_order(int64_t n_elems, const float *depths, const int32_t *tiles, int32_t *elems_out, int32_t *n_sorted_dev,
                   void *workspace, uint64_t workspace_bytes, cudaStream_t s, bool stats_ready) {
    RS_CHECK(n_elems >= 0 && n_elems < ((int64_t)1 << 30), "rs_depth_order: bad element count (limit 2^30)");
    RS_CHECK(n_sorted_dev != nullptr, "rs_depth_order: n_sorted_dev is required");
    if (n_elems == 0) {
        RS_CUDA(cudaMemsetAsync(n_sorted_dev, 0, sizeof(int32_t), s));
        return 0;
    }
    const DordLayout L = dord_layout(n_elems);
    RS_CHECK(depths && tiles && elems_out && workspace && workspace_bytes >= L.total,
             "rs_depth_order: null pointer or workspace too small (%llu < %llu)", (unsigned long long)workspace_bytes,
             (unsigned long long)L.total);
    char *w = reinterpret_cast<char *>(workspace);
    DordHeader *hdr = reinterpret_cast<DordHeader *>(w + L.hdr);
    unsigned int *counts = reinterpret_cast<unsigned int *>(w + L.counts);
    unsigned int *state = reinterpret_cast<unsigned int *>(w + L.state);
    unsigned int *group_start = reinterpret_cast<unsigned int *>(w + L.group_start);
    unsigned long long *comp = reinterpret_cast<unsigned long long *>(w + L.comp);
    unsigned long long *comp_alt = reinterpret_cast<unsigned long long *>(w + L.comp_alt);
    const int bits = dord_bucket_bits(n_elems);
    const int sms = rs_num_sms();
    const int grid = (int)min((int64_t)sms * 8, (n_elems + DORD_THREADS - 1) / DORD_THREADS);
    if (!stats_ready) { // else: cleared by rs_depth_order_prepare, statistics accumulated by the projection kernel
        if (int e = rs_depth_order_prepare(workspace, n_elems, s))
            return e;
        rs_dord_minmax_kernel<<<grid, DORD_THREADS, 0, s>>>(n_elems, depths, tiles, hdr);
        RS_LAUNCH_CHECK("rs_dord_minmax_kernel");
    }
    rs_dord_count_kernel<<<grid, DORD_THREADS, 0, s>>>(n_elems, depths, tiles, hdr, counts, bits);
    RS_LAUNCH_CHECK("rs_dord_count_kernel");
    rs_dord_scan_kernel<<<(1 << bits) / DORD_SCAN_PER_CTA, DORD_SCAN_THREADS, 0, s>>>(hdr, counts, state, group_start,
                                                                                      n_sorted_dev, bits);
    RS_LAUNCH_CHECK("rs_dord_scan_kernel");
    rs_dord_scatter_kernel<<<grid, DORD_THREADS, 0, s>>>(n_elems, depths, tiles, hdr, counts, comp);
    RS_LAUNCH_CHECK("rs_dord_scatter_kernel");
    const int sort_grid = (int)min((int64_t)sms * 8, n_elems / DORD_TARGET + 1);
    rs_dord_sort_kernel<<<sort_grid, DORD_THREADS, 0, s>>>(hdr, group_start, comp, comp_alt, elems_out);
    RS_LAUNCH_CHECK("rs_dord_sort_kernel");
    return 0;
}
