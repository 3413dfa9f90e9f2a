// rs_exchange_*: the one exchange step of the Gaussian-sharded render (BASELINE config c5) over NVLink peer memory.
// Replaces the all-to-alls of gsplat/rendering.py:527-611 (gsplat/distributed.py:10-257: a count exchange, then one NCCL
// all-to-all per attribute list) for the packed rows of rs_project_packed_fwd.
//
// Every rank owns ONE receive allocation (cudaMalloc, exported with cudaIpc*, mapped by every peer): a control block
// followed by the operator-layout arrays of the rows it will composite.  One kernel per rank and frame then
//   (A) publishes how many rows it holds for every destination into every peer's count table (+ a release flag),
//   (B) waits for the W count vectors, so every rank knows the full W x W matrix and therefore the exact first row of
//       its block inside every destination -- rows land compact, in (source rank, camera, Gaussian) order, exactly where
//       the reference's all-to-all puts them,
//   (C) stores its rows straight into the peers' arrays (camera ids made local, Gaussian ids made global, opacity x
//       compensation and the colour row gathered on the way: no cat / split / index / dtype-conversion passes), and
//   (D) the last CTA to finish raises this rank's data flag on every peer.
// rs_exchange_wait (one tiny kernel) holds the consumer stream until all W data flags of the frame are up.  No NCCL call,
// no host round trip inside the exchange; spins are bounded (RS_EXCHANGE_TIMEOUT_CYCLES) and report through the control
// block instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

// Spin limit of the flag waits: rs_exchange_args.timeout_ms, default RS_EXCHANGE_DEFAULT_TIMEOUT_MS (a rank may
// legitimately lag by a checkpoint or a GC pause; NCCL's own watchdog default is minutes).  clock64() ticks at the SM
// clock, taken as <= 2 GHz.
#define RS_EXCHANGE_DEFAULT_TIMEOUT_MS 60000
#define RS_EXCHANGE_CYCLES_PER_MS 2000000ll

struct ExchangeCtl {
    unsigned int cnt_flag[2][RS_EXCHANGE_MAX_WORLD];  // epoch of the count vector of source s (double-buffered by parity)
    unsigned int data_flag[RS_EXCHANGE_MAX_WORLD];    // epoch of the last complete row block of source s
    int counts[2][RS_EXCHANGE_MAX_WORLD][RS_EXCHANGE_MAX_WORLD]; // [parity][source][destination] rows
    unsigned int done_ctas;                           // local: CTAs of the running push kernel that finished
    unsigned int error;                               // local: 1 = spin timed out, 2 = a block did not fit
    unsigned int fail_epoch[RS_EXCHANGE_MAX_WORLD];   // epoch in which source s gave up waiting (it then sent NO rows)
};
static_assert(sizeof(ExchangeCtl) <= RS_EXCHANGE_CTL_BYTES, "control block too large");

static inline uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }

extern "C" int rs_exchange_layout(int64_t capacity, int32_t channels, uint64_t offsets[RS_EXCHANGE_COLUMNS + 1]) {
    RS_CHECK(capacity >= 0 && channels >= 1 && channels <= RS_MAX_CHANNELS, "rs_exchange_layout: bad capacity / channels");
    const uint64_t cap = (uint64_t)capacity;
    const uint64_t width[RS_EXCHANGE_COLUMNS] = {8, 4, 12, 4, 4ull * channels, 8, 8, 8};
    uint64_t off = RS_EXCHANGE_CTL_BYTES;
    for (int i = 0; i < RS_EXCHANGE_COLUMNS; ++i) {
        offsets[i] = off;
        off = align256(off + cap * width[i]);
    }
    offsets[RS_EXCHANGE_COLUMNS] = off;
    return 0;
}

extern "C" uint64_t rs_exchange_bytes(int64_t capacity, int32_t channels) {
    uint64_t offsets[RS_EXCHANGE_COLUMNS + 1];
    if (rs_exchange_layout(capacity, channels, offsets) != 0)
        return 0;
    return offsets[RS_EXCHANGE_COLUMNS];
}

// ---- peer memory (legacy CUDA IPC: plain cudaMalloc allocations, one per rank) -------------------------------------------
extern "C" int rs_peer_alloc(uint64_t bytes, void **ptr) {
    RS_CHECK(ptr != nullptr && bytes > 0, "rs_peer_alloc: bad arguments");
    RS_CUDA(cudaMalloc(ptr, bytes));
    RS_CUDA(cudaMemset(*ptr, 0, bytes));
    RS_CUDA(cudaDeviceSynchronize());
    return 0;
}
extern "C" int rs_peer_free(void *ptr) {
    if (ptr != nullptr)
        RS_CUDA(cudaFree(ptr));
    return 0;
}
extern "C" int rs_peer_export(void *ptr, uint8_t handle[RS_PEER_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == RS_PEER_HANDLE_BYTES, "IPC handle size");
    RS_CHECK(ptr != nullptr && handle != nullptr, "rs_peer_export: null argument");
    cudaIpcMemHandle_t h;
    RS_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, sizeof(h));
    return 0;
}
extern "C" int rs_peer_open(const uint8_t handle[RS_PEER_HANDLE_BYTES], void **ptr) {
    RS_CHECK(ptr != nullptr && handle != nullptr, "rs_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    RS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int rs_peer_close(void *ptr) {
    if (ptr != nullptr)
        RS_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

// ---- device side ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct ExchangeCols {
    float *means2d, *depths, *conics, *opacities, *colors;
    int32_t *radii;
    long long *camera_ids, *gaussian_ids;
};
__device__ __forceinline__ ExchangeCols exchange_cols(char *base, const unsigned long long *off) {
    ExchangeCols c;
    c.means2d = (float *)(base + off[0]);
    c.depths = (float *)(base + off[1]);
    c.conics = (float *)(base + off[2]);
    c.opacities = (float *)(base + off[3]);
    c.colors = (float *)(base + off[4]);
    c.radii = (int32_t *)(base + off[5]);
    c.camera_ids = (long long *)(base + off[6]);
    c.gaussian_ids = (long long *)(base + off[7]);
    return c;
}

struct ExchangeLayout {
    unsigned long long off[RS_EXCHANGE_COLUMNS];
};

// waits until flag[s] has reached `epoch` for every s < world (one lane per source); returns false on timeout
__device__ __forceinline__ bool wait_flags(const unsigned int *flags, int world, unsigned int epoch, int timeout_ms) {
    bool ok = true;
    const long long limit = (long long)(timeout_ms > 0 ? timeout_ms : RS_EXCHANGE_DEFAULT_TIMEOUT_MS) * RS_EXCHANGE_CYCLES_PER_MS;
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while ((int)(ld_sys(flags + threadIdx.x) - epoch) < 0) {
            if (clock64() - t0 > limit) {
                ok = false;
                break;
            }
            __nanosleep(64);
        }
    }
    return __syncthreads_and(ok);
}

__global__ void __launch_bounds__(256) rs_exchange_push_kernel(const rs_exchange_args a, const ExchangeLayout lay) {
    __shared__ long long dst_base[RS_EXCHANGE_MAX_WORLD]; // first row of this rank's block inside destination d
    __shared__ int src_lo[RS_EXCHANGE_MAX_WORLD + 1];      // local row range held for destination d
    const int W = a.world, r = a.rank, Cl = a.cameras_per_rank;
    const unsigned int par = a.epoch & 1u;
    char *const *peers = (char *const *)a.peer_base;
    ExchangeCtl *mine = (ExchangeCtl *)peers[r];

    // (A) publish my count vector to every peer, then the flag that releases it
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < W * W) {
            const int p = threadIdx.x / W, d = threadIdx.x % W;
            ((ExchangeCtl *)peers[p])->counts[par][r][d] = a.indptr[(d + 1) * Cl] - a.indptr[d * Cl];
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < W)
            st_sys(&((ExchangeCtl *)peers[threadIdx.x])->cnt_flag[par][r], a.epoch);
    }
    // (B) all W vectors -> placement of my blocks
    // A rank that gives up here must not place rows from a stale count matrix: it sends NOTHING this epoch (dst_base = -1
    // below), tells every peer so (fail_epoch, read by rs_exchange_wait on every rank, so all ranks raise together), and
    // still raises its data flag at the end so that no peer waits for it forever.
    const bool counts_ok = wait_flags(mine->cnt_flag[par], W, a.epoch, a.timeout_ms);
    if (!counts_ok) {
        if (threadIdx.x == 0)
            mine->error = 1u;
        if (blockIdx.x == 0 && (int)threadIdx.x < W)
            st_sys(&((ExchangeCtl *)peers[threadIdx.x])->fail_epoch[r], a.epoch);
    }
    if ((int)threadIdx.x <= W)
        src_lo[threadIdx.x] = a.indptr[min((int)threadIdx.x, W) * Cl];
    if ((int)threadIdx.x < W) {
        const int d = threadIdx.x;
        long long base = 0, total = 0;
        for (int s = 0; s < W; ++s) {
            const int c = *((volatile int *)&mine->counts[par][s][d]);
            if (s < r)
                base += c;
            total += c;
        }
        if (total > a.capacity) { // the block does not fit: nothing is written, the host regrows from the same matrix
            base = -1;
            mine->error = 2u;
        }
        dst_base[d] = counts_ok ? base : -1;
    }
    __syncthreads();

    // (C) rows -> peers, in blocks of 256 consecutive rows of one destination.  Every store instruction of a warp covers
    // one contiguous span of the destination column (word-granular copies of the direct columns, gathered columns formed
    // on the way), so the NVLink write packets are full: a row-per-thread loop would emit 12-byte fragments.
    __shared__ int blk_lo[RS_EXCHANGE_MAX_WORLD + 1]; // first block index of destination d in the flattened block list
    __shared__ long long gid_s[256];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int d = 0; d < W; ++d) {
            blk_lo[d] = acc;
            if (dst_base[d] >= 0)
                acc += (src_lo[d + 1] - src_lo[d] + 255) >> 8;
        }
        blk_lo[W] = acc;
    }
    __syncthreads();
    const int D = a.channels, n_blocks = blk_lo[W];
    const int t = threadIdx.x;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        int d = 0;
        while (blk >= blk_lo[d + 1])
            ++d;
        const long long row0 = (long long)src_lo[d] + ((long long)(blk - blk_lo[d]) << 8);
        const int n = (int)min(256ll, (long long)src_lo[d + 1] - row0);
        const long long out0 = dst_base[d] + ((long long)(blk - blk_lo[d]) << 8);
        const ExchangeCols c = exchange_cols(peers[d], lay.off);
        long long gid = 0;
        if (t < n) {
            gid = a.gaussian_ids[row0 + t];
            gid_s[t] = gid;
        }
        __syncthreads();
        for (int i = t; i < n * 2; i += 256) {
            c.means2d[out0 * 2 + i] = a.means2d[row0 * 2 + i];
            c.radii[out0 * 2 + i] = a.radii[row0 * 2 + i];
        }
        for (int i = t; i < n * 3; i += 256)
            c.conics[out0 * 3 + i] = a.conics[row0 * 3 + i];
        for (int i = t; i < n * D; i += 256) {
            const int rr = i / D, k = i - rr * D;
            c.colors[out0 * D + i] = a.colors_per_row ? a.colors[row0 * D + i] : a.colors[gid_s[rr] * D + k];
        }
        if (t < n) {
            c.depths[out0 + t] = a.depths[row0 + t];
            float op = a.opacities_per_row ? a.opacities[row0 + t] : a.opacities[gid];
            if (a.compensations != nullptr)
                op *= a.compensations[row0 + t];
            c.opacities[out0 + t] = op;
            c.camera_ids[out0 + t] = a.camera_ids[row0 + t] - (long long)d * Cl;
            c.gaussian_ids[out0 + t] = gid + a.gaussian_base;
        }
        __syncthreads();
    }

    // (D) last CTA raises my data flag everywhere
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0)
        last = atomicAdd(&mine->done_ctas, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0)
            mine->done_ctas = 0u;
        __threadfence_system();
        if ((int)threadIdx.x < W)
            st_sys(&((ExchangeCtl *)peers[threadIdx.x])->data_flag[r], a.epoch);
    }
}

// Holds the stream until the rows of every source have landed; then totals[0] = rows received by this rank,
// totals[1] = the largest row count any rank receives (what `capacity` must hold; identical on every rank),
// totals[2] = error code of this frame (0 ok, 1 timeout, 2 capacity), totals[3] = flags still behind (diagnostics).
__global__ void rs_exchange_wait_kernel(const rs_exchange_args a, long long *totals) {
    char *const *peers = (char *const *)a.peer_base;
    ExchangeCtl *mine = (ExchangeCtl *)peers[a.rank];
    const unsigned int par = a.epoch & 1u;
    bool ok = wait_flags(mine->data_flag, a.world, a.epoch, a.timeout_ms);
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.world; ++s) // a source that gave up on the counts sent no rows: the frame is invalid everywhere
            if (ld_sys(&mine->fail_epoch[s]) == a.epoch)
                ok = false;
        // diagnostics for a timed-out frame: which sources' flags are behind (bit s = data flag, bit 16 + s = count flag)
        long long behind = 0;
        for (int s = 0; s < a.world; ++s) {
            if ((int)(ld_sys(&mine->data_flag[s]) - a.epoch) < 0)
                behind |= 1ll << s;
            if ((int)(ld_sys(&mine->cnt_flag[par][s]) - a.epoch) < 0)
                behind |= 1ll << (16 + s);
        }
        totals[3] = behind;
        long long worst = 0, got = 0;
        for (int d = 0; d < a.world; ++d) {
            long long t = 0;
            for (int s = 0; s < a.world; ++s)
                t += *((volatile int *)&mine->counts[par][s][d]);
            worst = t > worst ? t : worst;
            if (d == a.rank)
                got = t;
        }
        totals[0] = got;
        totals[1] = worst;
        unsigned int err = mine->error;
        if (!ok)
            err = 1u;
        totals[2] = err;
        mine->error = 0u;
    }
}

// ---- transposed exchange (backward) -----------------------------------------------------------------------------------------
// Gradients of the rows a rank RECEIVED travel back to the ranks that sent them: the transposed all-to-all of
// gsplat/distributed.py:243-248 (the backward of its differentiable all_to_all), again as direct stores into peer memory.
// Every rank keeps the W x W row-count matrix of the forward exchange (rs_exchange_read_counts), so no count handshake is
// needed: the block received from source s -- rows [sum_{s'<s} counts[s'][r], +counts[s][r]) here -- goes to rows
// [sum_{d<r} counts[s][d], ...) of s's gradient arrays, i.e. exactly where s's packed rows for destination r sit.
// Each rank owns a second receive allocation for this (same layout function; only the five float columns are used).
__global__ void rs_exchange_counts_kernel(const rs_exchange_args a, int32_t *out) {
    char *const *peers = (char *const *)a.peer_base;
    const ExchangeCtl *mine = (const ExchangeCtl *)peers[a.rank];
    const unsigned int par = a.epoch & 1u;
    const int W = a.world;
    for (int i = threadIdx.x; i < W * W; i += blockDim.x)
        out[i] = *((volatile const int *)&mine->counts[par][i / W][i % W]);
}

__global__ void __launch_bounds__(256) rs_exchange_push_grad_kernel(const rs_exchange_grad_args a, const ExchangeLayout lay) {
    __shared__ long long src_lo[RS_EXCHANGE_MAX_WORLD + 1]; // my received rows of source s start here
    __shared__ long long dst_lo[RS_EXCHANGE_MAX_WORLD];     // ... and go to this row of s's gradient arrays
    __shared__ int blk_lo[RS_EXCHANGE_MAX_WORLD + 1];
    const int W = a.world, r = a.rank;
    char *const *peers = (char *const *)a.peer_base;
    ExchangeCtl *mine = (ExchangeCtl *)peers[r];
    if (threadIdx.x == 0) {
        long long acc = 0;
        int blocks = 0;
        for (int s = 0; s < W; ++s) {
            src_lo[s] = acc;
            blk_lo[s] = blocks;
            const long long n = a.counts[s * W + r];
            acc += n;
            blocks += (int)((n + 255) >> 8);
            long long before = 0;
            for (int d = 0; d < r; ++d)
                before += a.counts[s * W + d];
            dst_lo[s] = before;
        }
        src_lo[W] = acc;
        blk_lo[W] = blocks;
    }
    __syncthreads();
    const int D = a.channels, n_blocks = blk_lo[W];
    const int t = threadIdx.x;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        int s = 0;
        while (blk >= blk_lo[s + 1])
            ++s;
        const long long off = (long long)(blk - blk_lo[s]) << 8;
        const long long row0 = src_lo[s] + off;
        const int n = (int)min(256ll, src_lo[s + 1] - row0);
        const long long out0 = dst_lo[s] + off;
        if (out0 + n > a.capacity) { // cannot happen when the host sized the arrays from the same matrix; never write outside
            if (t == 0)
                mine->error = 2u;
            continue;
        }
        const ExchangeCols c = exchange_cols(peers[s], lay.off);
        for (int i = t; i < n * 2; i += 256)
            c.means2d[out0 * 2 + i] = a.v_means2d[row0 * 2 + i];
        for (int i = t; i < n * 3; i += 256)
            c.conics[out0 * 3 + i] = a.v_conics[row0 * 3 + i];
        for (int i = t; i < n * D; i += 256)
            c.colors[out0 * D + i] = a.v_colors[row0 * D + i];
        if (t < n) {
            c.depths[out0 + t] = a.v_depths[row0 + t];
            c.opacities[out0 + t] = a.v_opacities[row0 + t];
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0)
        last = atomicAdd(&mine->done_ctas, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0)
            mine->done_ctas = 0u;
        __threadfence_system();
        if ((int)threadIdx.x < W)
            st_sys(&((ExchangeCtl *)peers[threadIdx.x])->data_flag[r], a.epoch);
    }
}

__global__ void rs_exchange_wait_grad_kernel(const rs_exchange_grad_args a, long long *status) {
    char *const *peers = (char *const *)a.peer_base;
    ExchangeCtl *mine = (ExchangeCtl *)peers[a.rank];
    const bool ok = wait_flags(mine->data_flag, a.world, a.epoch, a.timeout_ms);
    if (threadIdx.x == 0) {
        long long behind = 0;
        for (int s = 0; s < a.world; ++s)
            if ((int)(ld_sys(&mine->data_flag[s]) - a.epoch) < 0)
                behind |= 1ll << s;
        unsigned int err = mine->error;
        if (!ok)
            err = 1u;
        status[0] = err;
        status[1] = behind;
        mine->error = 0u;
    }
}

// Rows [received, capacity) of this rank's receive arrays still hold an earlier frame.  rs_exchange_seal zeroes their radii,
// which makes them invisible to everything downstream (zero tiles), so that a caller can run tile binning and compositing over
// the whole CAPACITY with device-side counts only -- no host read of the row count between the exchange and the render.
__global__ void __launch_bounds__(256)
rs_exchange_seal_kernel(const rs_exchange_args a, const ExchangeLayout lay, const long long *totals) {
    char *const *peers = (char *const *)a.peer_base;
    int2 *radii = reinterpret_cast<int2 *>(peers[a.rank] + lay.off[5]);
    const long long got = totals[2] == 0 ? min(totals[0], (long long)a.capacity) : 0; // failed epoch: nothing is valid
    for (long long i = got + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.capacity; i += (long long)gridDim.x * blockDim.x)
        radii[i] = make_int2(0, 0);
}

static int exchange_check(const rs_exchange_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->world >= 1 && a->world <= RS_EXCHANGE_MAX_WORLD && a->rank >= 0 && a->rank < a->world,
             "%s: bad world / rank (%d / %d)", who, a->world, a->rank);
    RS_CHECK(a->cameras_per_rank >= 1 && a->channels >= 1 && a->channels <= RS_MAX_CHANNELS && a->capacity >= 0,
             "%s: bad cameras_per_rank / channels / capacity", who);
    RS_CHECK(a->peer_base != nullptr && a->epoch != 0u, "%s: peer table missing or epoch 0", who);
    return 0;
}

extern "C" int rs_exchange_push(const rs_exchange_args *a, rs_stream_t stream) {
    if (int rc = exchange_check(a, "rs_exchange_push"))
        return rc;
    RS_CHECK(a->indptr != nullptr && a->nnz >= 0, "rs_exchange_push: indptr / nnz missing");
    RS_CHECK(a->nnz == 0 || (a->camera_ids && a->gaussian_ids && a->radii && a->means2d && a->depths && a->conics &&
                             a->opacities && a->colors),
             "rs_exchange_push: null row pointer");
    ExchangeLayout lay;
    uint64_t off[RS_EXCHANGE_COLUMNS + 1];
    if (int rc = rs_exchange_layout(a->capacity, a->channels, off))
        return rc;
    for (int i = 0; i < RS_EXCHANGE_COLUMNS; ++i)
        lay.off[i] = off[i];
    // enough CTAs to keep the NVLink / HBM stores of every SM in flight (ncu at 2 per SM: 25 % occupancy, 9 % issue, 1.2 TB/s of
    // local copies); the count wait only depends on block 0 of the OTHER ranks' kernels, so CTAs beyond the resident set are
    // harmless
    const int grid = rs_num_sms() * 4;
    rs_exchange_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a, lay);
    RS_LAUNCH_CHECK("rs_exchange_push_kernel");
    return 0;
}

extern "C" int rs_exchange_wait(const rs_exchange_args *a, int64_t *totals_dev, rs_stream_t stream) {
    if (int rc = exchange_check(a, "rs_exchange_wait"))
        return rc;
    RS_CHECK(totals_dev != nullptr, "rs_exchange_wait: totals_dev is required");
    rs_exchange_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a, (long long *)totals_dev);
    RS_LAUNCH_CHECK("rs_exchange_wait_kernel");
    return 0;
}

extern "C" int rs_exchange_read_counts(const rs_exchange_args *a, int32_t *counts_dev, rs_stream_t stream) {
    if (int rc = exchange_check(a, "rs_exchange_read_counts"))
        return rc;
    RS_CHECK(counts_dev != nullptr, "rs_exchange_read_counts: counts_dev is required");
    rs_exchange_counts_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*a, counts_dev);
    RS_LAUNCH_CHECK("rs_exchange_counts_kernel");
    return 0;
}

static int exchange_grad_check(const rs_exchange_grad_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->world >= 1 && a->world <= RS_EXCHANGE_MAX_WORLD && a->rank >= 0 && a->rank < a->world,
             "%s: bad world / rank (%d / %d)", who, a->world, a->rank);
    RS_CHECK(a->channels >= 1 && a->channels <= RS_MAX_CHANNELS && a->capacity >= 0, "%s: bad channels / capacity", who);
    RS_CHECK(a->peer_base != nullptr && a->epoch != 0u && a->counts != nullptr, "%s: peer table / counts missing or epoch 0", who);
    return 0;
}

extern "C" int rs_exchange_push_grad(const rs_exchange_grad_args *a, rs_stream_t stream) {
    if (int rc = exchange_grad_check(a, "rs_exchange_push_grad"))
        return rc;
    RS_CHECK(a->v_means2d && a->v_depths && a->v_conics && a->v_opacities && a->v_colors, "rs_exchange_push_grad: null gradient pointer");
    ExchangeLayout lay;
    uint64_t off[RS_EXCHANGE_COLUMNS + 1];
    if (int rc = rs_exchange_layout(a->capacity, a->channels, off))
        return rc;
    for (int i = 0; i < RS_EXCHANGE_COLUMNS; ++i)
        lay.off[i] = off[i];
    rs_exchange_push_grad_kernel<<<rs_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(*a, lay);
    RS_LAUNCH_CHECK("rs_exchange_push_grad_kernel");
    return 0;
}

extern "C" int rs_exchange_wait_grad(const rs_exchange_grad_args *a, int64_t *status_dev, rs_stream_t stream) {
    if (int rc = exchange_grad_check(a, "rs_exchange_wait_grad"))
        return rc;
    RS_CHECK(status_dev != nullptr, "rs_exchange_wait_grad: status_dev is required");
    rs_exchange_wait_grad_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a, (long long *)status_dev);
    RS_LAUNCH_CHECK("rs_exchange_wait_grad_kernel");
    return 0;
}

extern "C" int rs_exchange_seal(const rs_exchange_args *a, const int64_t *totals_dev, rs_stream_t stream) {
    if (int rc = exchange_check(a, "rs_exchange_seal"))
        return rc;
    RS_CHECK(totals_dev != nullptr, "rs_exchange_seal: totals_dev (from rs_exchange_wait) is required");
    if (a->capacity == 0)
        return 0;
    ExchangeLayout lay;
    uint64_t off[RS_EXCHANGE_COLUMNS + 1];
    if (int rc = rs_exchange_layout(a->capacity, a->channels, off))
        return rc;
    for (int i = 0; i < RS_EXCHANGE_COLUMNS; ++i)
        lay.off[i] = off[i];
    const int grid = (int)min((int64_t)rs_num_sms() * 4, (a->capacity + 255) / 256);
    rs_exchange_seal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a, lay, (const long long *)totals_dev);
    RS_LAUNCH_CHECK("rs_exchange_seal_kernel");
    return 0;
}
