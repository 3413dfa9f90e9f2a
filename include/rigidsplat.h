/*
 * rigidsplat.h -- C ABI of librigidsplat.so: the B200-native (sm_100a) animate -> project -> tile-sort ->
 * composite hot path of JTStephens18/3DGS_rigidbody (a gsplat 1.5.3 fork).
 *
 * This header is the drop-in boundary.  Every entry point replaces one operator that the reference binds
 * through its pybind module `_C` (gsplat/cuda/ext.cpp:6-104, C++ signatures in gsplat/cuda/include/Ops.h);
 * the file:line each one stands in for is cited on the declaration.  Conventions:
 *   - plain C: POD structs, raw DEVICE pointers (float32 / int32 / int64), sizes, an explicit CUDA stream
 *     (`void*` = cudaStream_t; NULL = legacy default stream).  No torch / ATen types.
 *   - inputs are borrowed, must be contiguous and live on the current device; outputs are caller-allocated
 *     (the reference allocates them in its host launchers, e.g. csrc/Projection.cpp:144-162; our Python shim
 *     `3dgs_rigidbody_b200/_C.py` does that with torch and passes the pointers here).
 *   - every function returns 0 on success, non-zero on failure; `rs_last_error()` then holds a message
 *     (thread-local).  Launch errors are checked (the reference does not check them at all).
 *   - nothing here synchronises the stream or the device, except rs_isect_count_total() and rs_peer_alloc().
 *   - optional pointers may be NULL where marked "optional".
 *   - every call acts on the CURRENT device (cudaSetDevice is the caller's business); per-device state (SM count, opt-in
 *     shared-memory sizes of the kernels) is kept per device, so one process may drive several GPUs.
 */
#ifndef RIGIDSPLAT_H_
#define RIGIDSPLAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_ABI_VERSION 18

/* gsplat/cuda/include/Common.h:46-51 (CameraModelType) */
enum { RS_PINHOLE = 0, RS_ORTHO = 1, RS_FISHEYE = 2, RS_FTHETA = 3 };

typedef void *rs_stream_t; /* cudaStream_t */

int rs_abi_version(void);
const char *rs_last_error(void);
/* sizeof() of every args struct, for binding self-checks: which = 0 project_fwd, 1 project_bwd, 2 isect,
 * 3 sort, 4 raster_fwd, 5 raster_bwd, 6 frame, 7 rigid, 8 isect_sorted, 9 sh, 10 project_packed_fwd, 11 exchange, 12 cgc,
 * 13 seghead, 14 exchange_grad.  Returns 0 for an unknown id. */
uint64_t rs_sizeof_args(int which);
/* number of kernels this library has launched in this process (all threads); bench.py reports its delta over the
 * timed region as `gpu_launches`. */
uint64_t rs_launch_count(void);
/* Measurement aid (bench.py `roofline`): per-kernel durations of everything the CALLING THREAD launches through this
 * library on `stream` between rs_profile_begin() and rs_profile_end().  One CUDA event is recorded on `stream` behind each
 * kernel launch; rs_profile_end() synchronises on the last one and returns, per kernel in launch order, its name (static
 * string) and the time since the previous event in milliseconds (kernel + launch gap, as it ran in stream order).  At most
 * 256 kernels are recorded per profile. */
int rs_profile_begin(rs_stream_t stream);
int rs_profile_end(int32_t max_kernels, float *ms /* [max_kernels] */, const char **names /* [max_kernels] */,
                   int32_t *n_kernels);

/* ------------------------------------------------------------------------------------------------------------
 * Rigid pose table.  Replaces main.py:183-228 (apply_transform) + main.py:173-181 (quat_multiply) +
 * gsplat/utils.py:109-134 (normalized_quat_to_rotmat): instead of cloning every splat tensor per body and
 * running ~15 torch ops, the per-body pose is consumed inside the projection kernels.
 *   mean'  = R_k (mean - center_k) + center_k + trans_k      (main.py:210-213, 222)
 *   quat'  = (q_k / |q_k|) (x) quat        wxyz Hamilton     (main.py:207, 219)
 * cluster_ids[g] = k in [0,K) selects the body; k < 0 ("background") leaves the Gaussian untouched.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const int32_t *cluster_ids; /* [N] optional (NULL = no rigid transform) */
    const float *body_quats;    /* [K,4] wxyz, need not be unit */
    const float *body_trans;    /* [K,3] */
    const float *body_centers;  /* [K,3] optional (NULL = rotate about the origin) */
    int32_t K;
    int32_t _pad;
} rs_rigid_t;

/* ------------------------------------------------------------------------------------------------------------
 * rs_project_fwd: replaces `projection_ewa_3dgs_fused_fwd` (Ops.h:42-64, csrc/Projection.cpp:104-189, kernel
 * csrc/ProjectionEWA3DGSFused.cu:15-212) with the rigid transform fused in front, and (optionally) the first pass
 * of `intersect_tile` (csrc/IntersectTile.cu:55-84) fused behind it.
 * Culled entries get radii = 0 and ZEROS in means2d/depths/conics (the reference leaves them uninitialised).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t B, C, N;            /* batches, cameras per batch, Gaussians per batch */
    int32_t image_width, image_height;
    int32_t camera_model;       /* RS_PINHOLE | RS_ORTHO | RS_FISHEYE */
    float eps2d, near_plane, far_plane, radius_clip;
    const float *means;         /* [B,N,3] */
    const float *covars;        /* [B,N,6] optional, exclusive with quats+scales */
    const float *quats;         /* [B,N,4] optional */
    const float *scales;        /* [B,N,3] optional */
    const float *opacities;     /* [B,N] optional */
    const float *viewmats;      /* [B,C,4,4] row-major world->camera */
    const float *Ks;            /* [B,C,3,3] */
    rs_rigid_t rigid;           /* applied when rigid.cluster_ids != NULL (ids shared by all batches) */
    int32_t *radii;             /* [B,C,N,2] out */
    float *means2d;             /* [B,C,N,2] out */
    float *depths;              /* [B,C,N]   out */
    float *conics;              /* [B,C,N,3] out */
    float *compensations;       /* [B,C,N]   out, optional (non-NULL <=> calc_compensations) */
    /* optional fused tile counting: tiles_per_gauss (+ tile geometry); block_sums additionally feeds rs_isect_scan /
     * rs_isect_emit (the unsorted path) and may be NULL when only rs_isect_sorted consumes the counts: */
    int32_t *tiles_per_gauss;   /* [B*C*N] out, optional */
    int32_t *block_sums;        /* [rs_isect_num_blocks(B*C*N)] out, optional */
    int32_t tile_size, tile_width, tile_height;
    int32_t _pad;
    /* optional: compositing records [B*C*N, 8] float = {x, y, opacity, conic a | conic b, conic c, cull limit, 0},
     * the staging format of rs_raster_fwd (pass them as rs_raster_fwd_args.records with records_ready = 1).  Needs
     * opacities; written for visible rows only. */
    float *records;
    /* optional: view-dependent colours evaluated in place (rendering.py:491-525 without the dirs / colours round trip):
     * sh_colors[b,c,g,:] = max(SH(sh_degree, mean' - camera origin, sh_coeffs[b,g]) + 0.5, 0) for visible rows, where
     * mean' is the rigidly moved mean (the coefficients are NOT rotated with the body, as in main.py:200-226). */
    const float *sh_coeffs;     /* [B,N,sh_K,3] optional */
    float *sh_colors;           /* [B,C,N,3] out, required with sh_coeffs */
    int32_t sh_degree, sh_K;
    /* optional (needs tiles_per_gauss): statistics of the depth bits of the rows with at least one tile, accumulated with
     * atomics into depth_stats[0] = max(~bits), [1] = max(bits), [2] = number of such rows ([3] unused); must be zeroed by the
     * caller.  They are the first step of the depth ordering of rs_isect_sorted (depth_stats_ready), which rs_render_frame
     * fuses this way. */
    uint32_t *depth_stats;
    /* optional (needs tiles_per_gauss and opacities): tight tile lists.  tile_footprints[row] = {mask lo, mask hi,
     * x0 | y0 << 16, w | h << 16}: the bounding rectangle of tiles the reference lists (IntersectTile.cu:60-93) and, when it
     * has at most 64 tiles, bit t of the 64-bit mask = "tile t of the rectangle (row-major) holds a pixel centre where the
     * splat can reach alpha >= 1/255" -- every other (tile, splat) pair is skipped by every pixel of the tile in
     * RasterizeToPixels3DGSFwd.cu:148-149, so dropping it changes no pixel.  tiles_per_gauss then counts the mask bits
     * (rectangles of more than 64 tiles keep all their tiles).  Consumed by rs_isect_sorted_args.tile_footprints; the lists
     * are a subset of the reference's, the images are bit-identical. */
    uint32_t *tile_footprints;  /* [B*C*N,4] out, 16-byte aligned */
} rs_project_fwd_args;
int rs_project_fwd(const rs_project_fwd_args *a, rs_stream_t stream);

/* rs_project_bwd: replaces `projection_ewa_3dgs_fused_bwd` (Ops.h:65-88, csrc/Projection.cpp:191-281, kernel
 * csrc/ProjectionEWA3DGSFused.cu:293-531).  v_* outputs must be ZERO-initialised by the caller (the reference
 * zero-inits them in Projection.cpp:239-250); results are accumulated with atomics.  With a rigid table the
 * gradients are chained back to the UNtransformed means / quats (v_mean = R_k^T v_mean', v_quat = conj(q_k) (x) v_quat'). */
typedef struct {
    int32_t B, C, N;
    int32_t image_width, image_height;
    int32_t camera_model;
    float eps2d;
    int32_t _pad;
    const float *means, *covars, *quats, *scales, *viewmats, *Ks;
    rs_rigid_t rigid;
    const int32_t *radii;        /* [B,C,N,2] fwd output */
    const float *conics;         /* [B,C,N,3] fwd output */
    const float *compensations;  /* optional */
    const float *v_means2d;      /* [B,C,N,2] */
    const float *v_depths;       /* [B,C,N] */
    const float *v_conics;       /* [B,C,N,3] */
    const float *v_compensations;/* optional */
    float *v_means;              /* [B,N,3] optional */
    float *v_covars;             /* [B,N,6] optional (when covars given) */
    float *v_quats;              /* [B,N,4] optional */
    float *v_scales;             /* [B,N,3] optional */
    float *v_viewmats;           /* [B,C,4,4] optional */
    /* packed (COO) rows, replaces `projection_ewa_3dgs_packed_bwd` (Ops.h:125-151, csrc/Projection.cpp:415-547, kernel
     * csrc/ProjectionEWA3DGSPacked.cu:378-758): when gaussian_ids != NULL the per-row tensors (conics, compensations,
     * v_means2d, v_depths, v_conics, v_compensations) are [nnz, ...], row r belongs to (batch_ids[r], camera_ids[r],
     * gaussian_ids[r]), `radii` is ignored (every row is visible), and with sparse_grad != 0 the v_means / v_covars /
     * v_quats / v_scales outputs are [nnz, ...] rows instead of dense accumulators. */
    const int64_t *batch_ids, *camera_ids, *gaussian_ids;
    int64_t nnz;
    int32_t sparse_grad;
    int32_t _pad2;
} rs_project_bwd_args;
int rs_project_bwd(const rs_project_bwd_args *a, rs_stream_t stream);

/* rs_project_packed_fwd: replaces `projection_ewa_3dgs_packed_fwd` (Ops.h:98-124, csrc/Projection.cpp:283-413, kernels
 * csrc/ProjectionEWA3DGSPacked.cu:17-375).  The reference runs the projection TWICE (count pass, at::cumsum, a host
 * sync, write pass); here ONE pass projects every (image, Gaussian) pair and places the visible ones with a decoupled
 * look-back over 256-pair chunks taken in ticket order, so the rows come out in the reference's row-major
 * (batch, camera, gaussian) order without a second evaluation and without a host round trip inside the call.
 *   - `proj` holds the inputs of rs_project_fwd; its output pointers (radii, means2d, depths, conics, compensations and,
 *     if wanted, records / sh_colors / tiles_per_gauss) are PACKED rows [capacity, ...]; block_sums must be NULL.
 *   - rows beyond `capacity` are dropped but still counted: *nnz > capacity tells the caller to regrow.
 *   - indptr[i] = first row of image i (i = batch * C + camera), indptr[B*C] = nnz.
 *   - workspace: rs_project_packed_workspace_bytes(B, C, N) bytes of scratch (zeroed by the call itself). */
typedef struct {
    rs_project_fwd_args proj;
    int64_t capacity;
    int32_t *indptr;            /* [B*C+1] out, optional */
    int64_t *batch_ids;         /* [capacity] out */
    int64_t *camera_ids;        /* [capacity] out */
    int64_t *gaussian_ids;      /* [capacity] out */
    int64_t *nnz;               /* [1] device, out */
    void *workspace;
} rs_project_packed_fwd_args;
uint64_t rs_project_packed_workspace_bytes(int32_t B, int32_t C, int32_t N);
int rs_project_packed_fwd(const rs_project_packed_fwd_args *a, rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Gaussian-sharded scenes: the exchange of projected splats between the ranks of one NVLink domain.  Replaces the
 * all-to-alls of gsplat/rendering.py:527-611 on top of gsplat/distributed.py:10-257 (count exchange, then one NCCL
 * all-to-all per attribute list) for packed rows: every rank projects ITS Gaussians to ALL cameras
 * (rs_project_packed_fwd; rows ordered by camera, so the rows owed to one rank are contiguous), and ONE kernel per rank
 * stores them straight into the receive arrays of the ranks owning the cameras, over peer-mapped memory.  Rows land
 * compact and in (source rank, camera, Gaussian) order -- where the reference's all-to-all puts them -- with camera ids
 * made local and Gaussian ids made global.  See csrc/exchange.cu for the protocol.
 *
 * Receive allocation of a rank (identical layout on every rank): RS_EXCHANGE_CTL_BYTES of control block, then the
 * columns means2d f32[cap,2] | depths f32[cap] | conics f32[cap,3] | opacities f32[cap] | colors f32[cap,channels] |
 * radii i32[cap,2] | camera_ids i64[cap] | gaussian_ids i64[cap], each starting at offsets[i] (rs_exchange_layout).
 * It must be ZERO-filled when created (rs_peer_alloc does that) and mapped by every peer (rs_peer_export/open).
 * ------------------------------------------------------------------------------------------------------------ */
#define RS_EXCHANGE_MAX_WORLD 16
#define RS_EXCHANGE_COLUMNS 8
#define RS_EXCHANGE_CTL_BYTES 4096
#define RS_PEER_HANDLE_BYTES 64
int rs_exchange_layout(int64_t capacity, int32_t channels, uint64_t *offsets /* [RS_EXCHANGE_COLUMNS + 1]; last = total bytes */);
uint64_t rs_exchange_bytes(int64_t capacity, int32_t channels);
int rs_peer_alloc(uint64_t bytes, void **ptr);
int rs_peer_free(void *ptr);
int rs_peer_export(void *ptr, uint8_t *handle /* [RS_PEER_HANDLE_BYTES] out */);
int rs_peer_open(const uint8_t *handle, void **ptr);
int rs_peer_close(void *ptr);
typedef struct {
    int32_t world, rank;
    int32_t cameras_per_rank;    /* camera c of the gathered list belongs to rank c / cameras_per_rank */
    int32_t channels;
    int64_t capacity;            /* rows of every rank's receive arrays */
    uint32_t epoch;              /* frame counter, > 0, the same on every rank, +1 per exchange */
    int32_t colors_per_row;      /* colors is [nnz,channels] (1) or per local Gaussian [N,channels] (0) */
    int32_t opacities_per_row;   /* opacities is [nnz] (1) or per local Gaussian [N] (0) */
    int32_t timeout_ms;          /* spin limit of the flag waits; 0 = default (60 s).  A rank that times out on the counts
                                  * sends no rows and marks the epoch as failed on EVERY rank (error 1 everywhere) */
    void *const *peer_base;      /* device array [world]: this rank's mapping of every rank's receive allocation */
    /* this rank's packed rows (outputs of rs_project_packed_fwd with B = 1): */
    const int32_t *indptr;       /* [world*cameras_per_rank + 1] */
    const int64_t *camera_ids, *gaussian_ids;
    const int32_t *radii;
    const float *means2d, *depths, *conics;
    const float *compensations;  /* [nnz] optional: multiplied into the opacity on the way */
    const float *opacities;
    const float *colors;
    int64_t gaussian_base;       /* global index of this rank's first Gaussian */
    int64_t nnz;                 /* rows this rank holds (= indptr[last]); with 0 the row pointers may be NULL */
} rs_exchange_args;
/* publish counts, place, store rows into the peers, raise the data flags (one kernel) */
int rs_exchange_push(const rs_exchange_args *a, rs_stream_t stream);
/* hold `stream` until every source's rows of this epoch have landed; totals_dev (device, int64[4]) = {rows received,
 * largest row count any rank receives (capacity needed, identical on all ranks), error: 0 ok | 1 timeout | 2 capacity
 * exceeded -- nothing was written for the overfull destination, diagnostics: bit s = source s's data flag is behind,
 * bit 16 + s = its count flag is behind} */
int rs_exchange_wait(const rs_exchange_args *a, int64_t *totals_dev, rs_stream_t stream);
/* after rs_exchange_wait: zero the radii of the rows [received, capacity) of this rank's receive arrays (stale rows of earlier
 * frames), so that tile binning / compositing can run over the whole capacity with device-side counts only (a sync-free
 * sharded frame: distributed.ShardedFrameRenderer) */
int rs_exchange_seal(const rs_exchange_args *a, const int64_t *totals_dev, rs_stream_t stream);

/* The transposed exchange (backward of the above; replaces the backward of the differentiable all_to_all of
 * gsplat/distributed.py:243-248): gradients of the rows this rank RECEIVED are stored straight into the gradient arrays of
 * the ranks that sent them, at the positions of their packed rows.  rs_exchange_read_counts keeps the forward exchange's
 * W x W row-count matrix (counts[s * world + d] = rows source s sent to destination d) so that no handshake is needed later.
 * Every rank owns a SECOND receive allocation for gradients (rs_exchange_layout(capacity >= its own row count, channels);
 * only the float columns means2d | depths | conics | opacities | colors are used), mapped by every peer. */
typedef struct {
    int32_t world, rank, channels;
    int32_t timeout_ms;          /* spin limit of rs_exchange_wait_grad, 0 = default */
    int64_t capacity;            /* rows of every rank's gradient arrays */
    uint32_t epoch;              /* backward counter, > 0, the same on every rank, +1 per transposed exchange */
    int32_t _pad;
    void *const *peer_base;      /* device array [world]: this rank's mapping of every rank's GRADIENT allocation */
    const int32_t *counts;       /* device [world * world], from rs_exchange_read_counts of the matching forward */
    const float *v_means2d, *v_depths, *v_conics, *v_opacities, *v_colors; /* gradients of the received rows, forward order */
} rs_exchange_grad_args;
int rs_exchange_read_counts(const rs_exchange_args *a, int32_t *counts_dev /* [world * world] device */, rs_stream_t stream);
int rs_exchange_push_grad(const rs_exchange_grad_args *a, rs_stream_t stream);
/* hold `stream` until every peer's gradient rows of this epoch have landed; status_dev (device int64[2]) = {error: 0 ok |
 * 1 timeout | 2 a block did not fit, sources whose flag is behind (bit s)} */
int rs_exchange_wait_grad(const rs_exchange_grad_args *a, int64_t *status_dev, rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tile intersection.  Replaces `intersect_tile` (Ops.h:186-198, csrc/Intersect.cpp:15-149, kernels
 * csrc/IntersectTile.cu:23-207 and the cub::DeviceRadixSort call at :296-339) and `intersect_offset`
 * (Ops.h:199-204, csrc/Intersect.cpp:151-168, csrc/IntersectTile.cu:209-292).
 *
 * Key format (csrc/IntersectTile.cu:95-108): image_id << (32 + tile_n_bits) | tile_id << 32 | bits(depth),
 * value = flatten index (image * N + gaussian, or the nnz row when packed).
 * The reference needs one host sync to size its outputs (Intersect.cpp:79-80).  Here the count and the emission
 * are separate calls so that a caller with a pre-sized workspace never syncs:
 *   rs_isect_count      tiles_per_gauss + per-block sums                 (pass 1)
 *   rs_isect_scan       exclusive scan of the block sums, writes *n_isects (device)
 *   rs_isect_count_total  copies *n_isects to the host (THE sync of the compat path)
 *   rs_isect_emit       unsorted keys/values                              (pass 2; load-balanced, coalesced)
 *   rs_radix_sort_pairs hand-written stable LSD radix sort of (u64 key, i32 value) over [0, end_bit)
 *   rs_isect_offsets    first-isect index per (image, tile)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t n_elems;             /* I*N, or nnz when packed */
    int32_t N;                   /* Gaussians per image (unpacked); ignored when image_ids != NULL */
    int32_t I;                   /* number of images */
    int32_t tile_size, tile_width, tile_height;
    const float *means2d;        /* [n_elems,2] */
    const int32_t *radii;        /* [n_elems,2] */
    const float *depths;         /* [n_elems] */
    const int64_t *image_ids;    /* [nnz] optional: packed mode */
    int32_t *tiles_per_gauss;    /* [n_elems] out (count) / in (emit) */
    int32_t *block_sums;         /* [rs_isect_num_blocks(n_elems) + 1] scratch: sums -> exclusive offsets */
    int32_t *n_isects;           /* [1] device, written by rs_isect_scan */
    int64_t *isect_ids;          /* [capacity] out (emit) */
    int32_t *flatten_ids;        /* [capacity] out (emit) */
    int64_t capacity;            /* emit writes nothing beyond it; *overflow set when n_isects > capacity */
    int32_t *overflow;           /* [1] device, optional */
} rs_isect_args;
int32_t rs_isect_num_blocks(int64_t n_elems);
int rs_isect_count(const rs_isect_args *a, rs_stream_t stream);
int rs_isect_scan(const rs_isect_args *a, rs_stream_t stream);
int rs_isect_count_total(const rs_isect_args *a, rs_stream_t stream, int64_t *n_isects_host);
int rs_isect_emit(const rs_isect_args *a, rs_stream_t stream);

/* rs_isect_sorted: emission + sort + offsets in one sync-free call, producing exactly what `intersect_tile(sort=True)`
 * + `intersect_offset` produce (same key order, same tie order) without ever sorting 64-bit keys: the (depth, flatten
 * index) pairs are ordered first, intersections are emitted in that order and only their (image | tile) bits are sorted
 * (see csrc/isect.cu).  Inputs: isect.{means2d,radii,depths,tiles_per_gauss (from rs_isect_count or rs_project_fwd),
 * image_ids}; outputs: isect.flatten_ids [capacity] (sorted), isect.isect_ids [capacity] (sorted, optional),
 * tile_offsets (optional), *isect.n_isects and *isect.overflow (device).  isect.block_sums is not used. */
typedef struct {
    rs_isect_args isect;
    int32_t *tile_offsets;       /* [I,tile_height,tile_width] out, optional */
    void *workspace;             /* rs_isect_sorted_workspace_bytes(n_elems, capacity) bytes */
    uint64_t workspace_bytes;
    /* != 0: the depth statistics of the visible rows were already accumulated into rs_isect_sorted_depth_stats(workspace)
     * by the producer of the rows (rs_project_fwd_args.depth_stats) after rs_isect_sorted_prepare(); 0: computed here */
    int32_t depth_stats_ready;
    int32_t _pad;
    /* optional: emit only the tiles of rs_project_fwd_args.tile_footprints (isect.tiles_per_gauss must be the counts written
     * with them); NULL: every tile of the bounding rectangle, exactly the reference's lists */
    const uint32_t *tile_footprints;
} rs_isect_sorted_args;
/* for a caller that fuses the depth statistics into its projection: clear the ordering state of `workspace` (enqueued on
 * `stream`, BEFORE the kernel that accumulates the statistics) / where that kernel has to accumulate them */
/* Tile counts (isect.tiles_per_gauss) + tile footprints (format of rs_project_fwd_args.tile_footprints) of projected rows
 * that did not come out of rs_project_fwd -- e.g. the receive arrays of the splat exchange -- in one pass, for
 * rs_isect_sorted_args.tile_footprints.  conics / opacities ([n_elems,3] / [n_elems], per row) both given: tight lists;
 * both NULL: every tile of the bounding rectangle, exactly the reference's lists (the emission then needs one 16-byte
 * record per row instead of radii + means2d + count).  isect.block_sums, when given, receives the per-block sums as from
 * rs_isect_count (for rs_isect_scan / rs_isect_count_total). */
int rs_isect_footprints(const rs_isect_args *a, const float *conics, const float *opacities, uint32_t *tile_footprints,
                        rs_stream_t stream);
int rs_isect_sorted_prepare(void *workspace, int64_t n_elems, int64_t capacity, rs_stream_t stream);
uint32_t *rs_isect_sorted_depth_stats(void *workspace, int64_t n_elems, int64_t capacity);
uint64_t rs_isect_sorted_workspace_bytes(int64_t n_elems, int64_t capacity);
int rs_isect_sorted(const rs_isect_sorted_args *a, rs_stream_t stream);

typedef struct {
    int64_t n;                   /* number of pairs if n_dev == NULL, else capacity (grid sizing bound) */
    const int32_t *n_dev;        /* optional device count (sync-free path): only the first *n_dev pairs are sorted */
    int32_t begin_bit, end_bit;  /* sort on key bits [begin_bit, end_bit) */
    int64_t *keys_a, *keys_b;    /* double buffer; input in keys_a */
    int32_t *vals_a, *vals_b;    /* double buffer; input in vals_a */
    void *workspace;             /* rs_radix_sort_workspace_bytes(n) bytes */
    uint64_t workspace_bytes;
    int32_t *result_in_b;        /* host out: 1 if the sorted data ended up in keys_b/vals_b, else 0 */
} rs_sort_args;
uint64_t rs_radix_sort_workspace_bytes(int64_t n);
int rs_radix_sort_pairs(const rs_sort_args *a, rs_stream_t stream);
/* the same sort for 32-bit keys (keys_a / keys_b are uint32 arrays, end_bit <= 32): what the binning path runs on the
 * depth keys and on the (image | tile) keys */
int rs_radix_sort_pairs32(const rs_sort_args *a, rs_stream_t stream);

int rs_isect_offsets(const int64_t *isect_ids_sorted, int64_t n_isects, const int32_t *n_isects_dev /*optional*/,
                     int32_t I, int32_t tile_width, int32_t tile_height, int32_t *offsets /*[I*th*tw]*/,
                     rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Compositing.  rs_raster_fwd replaces `rasterize_to_pixels_3dgs_fwd` (Ops.h:223-238, csrc/Rasterization.cpp:20-115,
 * kernel csrc/RasterizeToPixels3DGSFwd.cu:17-187); rs_raster_bwd replaces `rasterize_to_pixels_3dgs_bwd`
 * (Ops.h:239-263, csrc/Rasterization.cpp:117-228, kernel csrc/RasterizeToPixels3DGSBwd.cu:15-276).
 * Any channel count 1..RS_MAX_CHANNELS is accepted directly (the reference pads to one of 19 template
 * instantiations in python, _wrapper.py:604-648).  tile_size must be 16 (the only value the reference exercises,
 * rendering.py:184-185).
 * ------------------------------------------------------------------------------------------------------------ */
#define RS_MAX_CHANNELS 513
typedef struct {
    int32_t I, N;                /* images; Gaussians per image (0 when packed) */
    int32_t channels;
    int32_t image_width, image_height, tile_size, tile_width, tile_height;
    int64_t n_isects;            /* used when n_isects_dev == NULL */
    const int32_t *n_isects_dev; /* optional device count */
    const float *means2d;        /* [I*N,2] or [nnz,2] */
    const float *conics;         /* [I*N,3] */
    const float *colors;         /* [I*N,channels] */
    const float *opacities;      /* [I*N] */
    const float *backgrounds;    /* [I,channels] optional */
    const uint8_t *masks;        /* [I,tile_height,tile_width] bool, optional */
    const int32_t *tile_offsets; /* [I,tile_height,tile_width] */
    const int32_t *flatten_ids;  /* [n_isects] */
    /* Broadcast without materialising [I,N,...] copies (rendering.py:446-448, 481-485 do torch.broadcast_to + contiguous):
     * when > 0 the colour / opacity row of flatten id g is g % attr_mod (i.e. shared by all images). 0 = per-image rows. */
    int32_t attr_mod_colors;
    int32_t attr_mod_opacities;
    float *render_colors;        /* [I,H,W,channels] out */
    float *render_alphas;        /* [I,H,W,1] out */
    int32_t *last_ids;           /* [I,H,W] out */
    /* Staging records [I*N or nnz, 8] float (caller-allocated scratch, 32 B per row, 16 B aligned).  With
     * records_ready == 0 rs_raster_fwd first packs them from means2d / conics / opacities; with records_ready == 1 they
     * were written by rs_project_fwd (frame path) and means2d / conics / opacities may be NULL. */
    float *records;
    int32_t records_ready;
    int32_t _pad;
    int64_t n_rows;              /* rows of means2d / conics (I*N, or nnz when packed); needed when records_ready == 0 */
    /* optional 8-bit frame [I,H,W,3] of the first three channels, quantised the way the reference's animation loop stores
     * a frame (main.py:140-171 save_rendered_image -> torchvision save_image: x * 255 + 0.5, clamp to [0, 255], truncate);
     * needs channels >= 3.  Written by the compositing epilogue next to the float image. */
    uint8_t *render_rgb8;
    /* optional device uint32, ZERO when the call is enqueued (and not shared with another launch in flight): work counter of
     * the persistent compositing kernel (tiles are then handed out dynamically instead of in a fixed stride).  NULL is fine. */
    uint32_t *tile_counter;
} rs_raster_fwd_args;
int rs_raster_fwd(const rs_raster_fwd_args *a, rs_stream_t stream);

typedef struct {
    rs_raster_fwd_args f;        /* forward inputs (render_colors unused; render_alphas/last_ids are inputs here).
                                  * f.records (+ f.n_rows, f.records_ready) as in rs_raster_fwd: with the scratch the splat
                                  * batches are staged through a shared-memory ring by a producer warp (the default
                                  * kernel); with f.records == NULL the barrier-per-batch kernel runs.  Same results. */
    const float *v_render_colors;/* [I,H,W,channels] */
    const float *v_render_alphas;/* [I,H,W,1] */
    float *v_means2d_abs;        /* [I*N,2] optional (absgrad); zero-initialised by the caller */
    float *v_means2d;            /* [I*N,2] zero-initialised by the caller */
    float *v_conics;             /* [I*N,3] zero-initialised */
    float *v_colors;             /* [I*N,channels] zero-initialised */
    float *v_opacities;          /* [I*N] zero-initialised */
} rs_raster_bwd_args;
int rs_raster_bwd(const rs_raster_bwd_args *a, rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Spherical harmonics.  rs_sh_fwd replaces `spherical_harmonics_fwd` (Ops.h:154-160, csrc/SphericalHarmonicsCUDA.cu:20-116,
 * :374-400), rs_sh_bwd `spherical_harmonics_bwd` (Ops.h:161-168, :118-372, :403-540): real SH up to degree 4 of the
 * normalised direction, coefficients [n, K, 3].  Rows with masks[i] == 0 produce zeros (the reference leaves them
 * uninitialised).  v_coeffs is fully written by rs_sh_bwd (rows beyond the used degree are zero).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t n;
    int32_t degree;              /* degrees_to_use, 0..4 */
    int32_t K;                   /* coefficient rows per element, >= (degree+1)^2 */
    const float *dirs;           /* [n,3], need not be unit */
    const float *coeffs;         /* [n,K,3] */
    const uint8_t *masks;        /* [n] bool, optional */
    float *colors;               /* [n,3] out (forward) */
    const float *v_colors;       /* [n,3] (backward) */
    float *v_coeffs;             /* [n,K,3] out (backward) */
    float *v_dirs;               /* [n,3] out (backward), optional */
} rs_sh_args;
int rs_sh_fwd(const rs_sh_args *a, rs_stream_t stream);
int rs_sh_bwd(const rs_sh_args *a, rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Contrastive clustering loss of the identity-feature training step (the consumer of the c3 compositing pass).
 * rs_cgc_fwd / rs_cgc_bwd replace `cgc_contrastive_clustering_loss` (examples/utils.py:828-904, called at
 * examples/simple_trainer.py:945-975 on the rendered [H,W,D] feature map) and its autograd: three passes over the
 * feature map forward, one more backward, instead of ~25 torch kernels each way.  The host derives, from the instance mask
 * alone, for every pixel its `member` cluster (index among the K valid foreground clusters, -1 otherwise) and its `target`
 * (same, except that -- a quirk of the reference reproduced on purpose, utils.py:878-883 -- background pixels are active
 * with the LAST foreground cluster as target when that cluster is valid), plus the per-cluster pixel counts.
 * ------------------------------------------------------------------------------------------------------------ */
#define RS_CGC_MAX_DIM 32
#define RS_CGC_MAX_CLUSTERS 128
typedef struct {
    int64_t P;                   /* pixels */
    int64_t A;                   /* active pixels (target >= 0), >= 1 */
    int32_t D;                   /* feature channels, 1..RS_CGC_MAX_DIM */
    int32_t K;                   /* valid clusters, 2..RS_CGC_MAX_CLUSTERS */
    float eps;                   /* floor of the temperatures (reference default 1e-6) */
    int32_t accumulate_grad;     /* rs_cgc_fwd: also accumulate what rs_cgc_bwd needs */
    const float *features;       /* [P,D] rendered feature map (un-normalised) */
    const int32_t *target;       /* [P] */
    const int32_t *member;       /* [P] */
    const float *n_member;       /* [K] member pixels per cluster (all >= min_cluster_size) */
    const float *n_active;       /* [K] active pixels per target */
    float *ws;                   /* rs_cgc_workspace_floats(K, D) floats; ws[last] = the loss after rs_cgc_fwd */
    const float *grad_loss;      /* [1] device, optional upstream gradient (default 1) */
    float *v_features;           /* [P,D] out (rs_cgc_bwd) */
} rs_cgc_args;
uint64_t rs_cgc_workspace_floats(int32_t K, int32_t D);
int rs_cgc_fwd(const rs_cgc_args *a, rs_stream_t stream);
int rs_cgc_bwd(const rs_cgc_args *a, rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Segmentation head of the identity-feature training step, fused: rs_seghead_fwd / rs_seghead_bwd replace
 * `torch.nn.Sequential(Linear(D, H), ReLU(), Linear(H, D))` applied to the per-Gaussian identity encodings
 * (examples/simple_trainer.py:442-446 builds it with D = identity_dim = 16, H = 64; :946-947 applies it to all N
 * encodings every step) and its autograd.  y = W2 relu(W1 x + b1) + b2 with the torch.nn.Linear weight layout
 * (W1 [H, D], W2 [D, H], row-major).  The [N, H] hidden layer is never written to HBM: forward streams x -> y, backward
 * recomputes the activations per row and accumulates the weight gradients as register-resident tile products.
 * Forward: any D <= 32, H <= 128.  Backward: D = 16, H = 64 (the reference's configuration).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t N;                   /* rows (Gaussians) */
    int32_t D, H;                /* feature / hidden width */
    const float *x;              /* [N, D] */
    const float *w1, *b1;        /* [H, D], [H] */
    const float *w2, *b2;        /* [D, H], [D] */
    float *y;                    /* [N, D] out (rs_seghead_fwd) */
    const float *v_y;            /* [N, D] (rs_seghead_bwd) */
    float *v_x;                  /* [N, D] out (rs_seghead_bwd), optional */
    float *v_w1, *v_b1;          /* [H, D], [H]  ACCUMULATED into: zero-initialised by the caller */
    float *v_w2, *v_b2;          /* [D, H], [D]  ACCUMULATED into */
} rs_seghead_args;
int rs_seghead_fwd(const rs_seghead_args *a, rs_stream_t stream);
int rs_seghead_bwd(const rs_seghead_args *a, rs_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * rs_render_frame: the whole per-frame hot path in one call with NO host synchronisation -- what the commented-out
 * animation loop of main.py:357-409 would run per frame (apply_transform per body + rasterization(), rendering.py:33-770,
 * packed=False, sh_degree=None).  All intermediates live in a caller-owned workspace sized by rs_frame_workspace_bytes();
 * if a frame produces more tile intersections than `max_isects`, *status (device int32[4]: {n_isects, overflow, 0, 0})
 * reports it and the surplus intersections are dropped (the caller re-renders with a larger workspace).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    rs_project_fwd_args proj;    /* B must be 1; tiles_per_gauss/block_sums are taken from the workspace when NULL */
    const float *colors;         /* [N,channels] (shared by all cameras) or [C,N,channels] if colors_per_camera;
                                  * ignored when proj.sh_coeffs is set (colours then come from the SH evaluation fused
                                  * into the projection, channels must be 3, proj.sh_colors is taken from the workspace) */
    int32_t channels;
    int32_t colors_per_camera;
    const float *backgrounds;    /* [C,channels] optional */
    int64_t max_isects;
    void *workspace;
    uint64_t workspace_bytes;
    float *render_colors;        /* [C,H,W,channels] out */
    float *render_alphas;        /* [C,H,W,1] out */
    int32_t *status;             /* device int32[4] out */
    /* optional exports of the sorted intersection data (`meta` of rendering.py:651-665); NULL to skip */
    int32_t *out_tile_offsets;   /* [C,tile_h,tile_w] */
    /* which part of the frame this call enqueues: 0 = all of it; RS_FRAME_BIN = rigid + projection + binning only;
     * RS_FRAME_COMPOSITE = compositing only (of a frame whose RS_FRAME_BIN part was enqueued with the same arguments).
     * Lets a caller put the short latency-bound binning kernels of the next frame on a high-priority stream so that they
     * are dispatched underneath the compositing of the previous one (FramePipeline). */
    int32_t stages;
    /* != 0: tight tile lists (rs_project_fwd_args.tile_footprints): the same image bit for bit from fewer intersections;
     * the exported lists (out_tile_offsets, rs_frame_workspace_ptr) are then a subset of the reference's.  0: the
     * reference's lists exactly. */
    int32_t tight_tiles;
    uint8_t *render_rgb8;        /* [C,H,W,3] out, optional: the frame as 8-bit RGB (see rs_raster_fwd_args.render_rgb8) */
} rs_frame_args;
#define RS_FRAME_BIN 1
#define RS_FRAME_COMPOSITE 2
uint64_t rs_frame_workspace_bytes(int32_t C, int32_t N, int32_t image_width, int32_t image_height, int32_t tile_size,
                                  int32_t channels, int64_t max_isects);
int rs_render_frame(const rs_frame_args *a, rs_stream_t stream);
/* Measurement aid: the same frame with CUDA events between its stages; synchronises the stream.  stage_ms (host, [4]) =
 * {rigid + projection, binning (sorts, emission, offsets), compositing, whole frame} in milliseconds. */
int rs_render_frame_timed(const rs_frame_args *a, rs_stream_t stream, float *stage_ms);
/* device pointers into a frame workspace, for tests and `meta`: which = 0 isect_ids(sorted), 1 flatten_ids(sorted),
 * 2 tile_offsets, 3 last_ids, 4 tiles_per_gauss.  Valid after rs_render_frame on the same workspace geometry. */
void *rs_frame_workspace_ptr(const rs_frame_args *a, int which);
/* The frame path never materialises 64-bit keys; this rebuilds the sorted isect ids of the last frame (which = 0 above)
 * from its offsets table, flatten ids and depths.  Enqueued on `stream`, no sync. */
int rs_frame_export_isect_ids(const rs_frame_args *a, rs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RIGIDSPLAT_H_ */
