// rs_render_frame: animate -> project -> tile-sort -> composite for one frame, no host synchronisation.
// This is what one iteration of the (commented-out) animation loop of main.py:357-409 does through
// apply_transform() + rasterization() (rendering.py:33-770, packed=False, sh_degree=None), minus the per-body tensor
// clones, the torch glue kernels and the `.item()` sync of csrc/Intersect.cpp:80.
#include <string.h>

#include "common.cuh"

int rs_isect_ids_from_offsets(const int32_t *offsets, const int32_t *flatten_ids, const float *depths, int64_t n_bound,
                              const int32_t *n_dev, int32_t I, int32_t tile_width, int32_t tile_height,
                              int64_t *isect_ids, rs_stream_t stream);

namespace {
struct FrameLayout {
    size_t radii, means2d, depths, conics, records, sh_colors, tiles_per_gauss, block_sums, isect_ids, flatten_ids, bin_ws, bin_ws_bytes,
        tile_offsets, last_ids, tile_counter, footprints, total;
};
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

FrameLayout make_layout(int32_t C, int32_t N, int32_t W, int32_t H, int32_t tile_size, int64_t max_isects) {
    FrameLayout L;
    const size_t E = (size_t)C * N;
    const size_t tw = (W + tile_size - 1) / tile_size, th = (H + tile_size - 1) / tile_size;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align256(o + bytes);
        return at;
    };
    L.radii = take(E * 2 * 4);
    L.means2d = take(E * 2 * 4);
    L.depths = take(E * 4);
    L.conics = take(E * 3 * 4);
    L.records = take(E * 8 * 4);
    L.sh_colors = take(E * 3 * 4); // only used when colours come from SH coefficients
    L.tiles_per_gauss = take(E * 4);
    L.block_sums = take(((size_t)rs_isect_num_blocks((int64_t)E) + 1) * 4);
    L.isect_ids = take((size_t)max_isects * 8); // only written by rs_frame_export_isect_ids
    L.flatten_ids = take((size_t)max_isects * 4);
    L.bin_ws_bytes = rs_isect_sorted_workspace_bytes((int64_t)E, max_isects);
    L.bin_ws = take(L.bin_ws_bytes);
    L.tile_offsets = take((size_t)C * tw * th * 4);
    L.last_ids = take((size_t)C * W * H * 4);
    L.tile_counter = take(256); // work counter of the persistent compositing kernel (zeroed per frame)
    L.footprints = take(E * 16); // tight tile lists (rs_frame_args.tight_tiles)
    L.total = o;
    return L;
}
} // namespace

extern "C" uint64_t rs_frame_workspace_bytes(int32_t C, int32_t N, int32_t image_width, int32_t image_height,
                                             int32_t tile_size, int32_t channels, int64_t max_isects) {
    (void)channels;
    if (C <= 0 || N < 0 || image_width <= 0 || image_height <= 0 || tile_size <= 0 || max_isects < 0)
        return 0;
    return make_layout(C, N, image_width, image_height, tile_size, max_isects).total;
}

extern "C" void *rs_frame_workspace_ptr(const rs_frame_args *a, int which) {
    if (a == nullptr || a->workspace == nullptr)
        return nullptr;
    const rs_project_fwd_args &p = a->proj;
    const FrameLayout L = make_layout(p.C, p.N, p.image_width, p.image_height, p.tile_size, a->max_isects);
    char *w = reinterpret_cast<char *>(a->workspace);
    switch (which) {
    case 0:
        return w + L.isect_ids;
    case 1:
        return w + L.flatten_ids;
    case 2:
        return w + L.tile_offsets;
    case 3:
        return w + L.last_ids;
    case 4:
        return w + L.tiles_per_gauss;
    case 5:
        return w + L.radii;
    case 6:
        return w + L.means2d;
    case 7:
        return w + L.depths;
    case 8:
        return w + L.conics;
    default:
        return nullptr;
    }
}

static void fill_isect(rs_isect_args &ia, const rs_project_fwd_args &p, const rs_frame_args *a, char *w,
                       const FrameLayout &L) {
    ia.n_elems = p.C * p.N;
    ia.N = p.N;
    ia.I = p.C;
    ia.tile_size = p.tile_size;
    ia.tile_width = (p.image_width + p.tile_size - 1) / p.tile_size;
    ia.tile_height = (p.image_height + p.tile_size - 1) / p.tile_size;
    ia.means2d = reinterpret_cast<const float *>(w + L.means2d);
    ia.radii = reinterpret_cast<const int32_t *>(w + L.radii);
    ia.depths = reinterpret_cast<const float *>(w + L.depths);
    ia.tiles_per_gauss = reinterpret_cast<int32_t *>(w + L.tiles_per_gauss);
    ia.n_isects = a->status;     // status[0]
    ia.overflow = a->status + 1; // status[1]
    ia.isect_ids = nullptr;      // 64-bit ids are not needed to composite; see rs_frame_export_isect_ids
    ia.flatten_ids = reinterpret_cast<int32_t *>(w + L.flatten_ids);
    ia.capacity = a->max_isects;
}

// Rebuilds the sorted 64-bit isect ids of the last frame rendered into this workspace (`meta["isect_ids"]`,
// rendering.py:651-665): key = (image | tile) << 32 | depth bits, recovered from the offsets table and flatten ids.
extern "C" int rs_frame_export_isect_ids(const rs_frame_args *a, rs_stream_t stream) {
    RS_CHECK(a != nullptr && a->workspace != nullptr && a->status != nullptr, "rs_frame_export_isect_ids: null args");
    const rs_project_fwd_args &p = a->proj;
    const FrameLayout L = make_layout(p.C, p.N, p.image_width, p.image_height, p.tile_size, a->max_isects);
    char *w = reinterpret_cast<char *>(a->workspace);
    const int tw = (p.image_width + p.tile_size - 1) / p.tile_size, th = (p.image_height + p.tile_size - 1) / p.tile_size;
    const int32_t *offsets = a->out_tile_offsets != nullptr ? a->out_tile_offsets : reinterpret_cast<const int32_t *>(w + L.tile_offsets);
    return rs_isect_ids_from_offsets(offsets, reinterpret_cast<const int32_t *>(w + L.flatten_ids),
                                     reinterpret_cast<const float *>(w + L.depths), a->max_isects, a->status, p.C, tw, th,
                                     reinterpret_cast<int64_t *>(w + L.isect_ids), stream);
}

static int render_frame_impl(const rs_frame_args *a, rs_stream_t stream, cudaEvent_t *ev);

extern "C" int rs_render_frame(const rs_frame_args *a, rs_stream_t stream) { return render_frame_impl(a, stream, nullptr); }

// Same frame with CUDA events between the stages; synchronises the stream and returns the stage times in milliseconds:
// stage_ms[0] rigid + projection (+ tile count, records), [1] depth-ordered binning (both sorts, emission, offsets),
// [2] compositing, [3] whole frame.  Measurement aid for bench.py; the product path is rs_render_frame.
extern "C" int rs_render_frame_timed(const rs_frame_args *a, rs_stream_t stream, float *stage_ms) {
    RS_CHECK(stage_ms != nullptr, "rs_render_frame_timed: null stage_ms");
    static thread_local cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i)
        if (ev[i] == nullptr)
            RS_CUDA(cudaEventCreate(&ev[i]));
    if (int e = render_frame_impl(a, stream, ev))
        return e;
    RS_CUDA(cudaEventSynchronize(ev[3]));
    for (int i = 0; i < 3; ++i)
        RS_CUDA(cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]));
    RS_CUDA(cudaEventElapsedTime(&stage_ms[3], ev[0], ev[3]));
    return 0;
}

static int render_frame_impl(const rs_frame_args *a, rs_stream_t stream, cudaEvent_t *ev) {
    RS_CHECK(a != nullptr, "rs_render_frame: null args");
    rs_project_fwd_args p = a->proj;
    RS_CHECK(p.B == 1, "rs_render_frame: B must be 1");
    RS_CHECK(p.C >= 1 && p.N >= 0, "rs_render_frame: bad sizes");
    RS_CHECK(p.tile_size == RS_TILE, "rs_render_frame: tile_size must be 16");
    RS_CHECK(a->channels >= 1 && a->channels <= RS_MAX_CHANNELS, "rs_render_frame: Unsupported number of color channels: %d",
             a->channels);
    RS_CHECK((a->colors || p.sh_coeffs) && a->render_colors && a->render_alphas && a->status && a->workspace,
             "rs_render_frame: null pointer");
    RS_CHECK(p.sh_coeffs == nullptr || a->channels == 3, "rs_render_frame: SH colours have 3 channels");
    RS_CHECK(p.opacities != nullptr, "rs_render_frame: opacities required");
    RS_CHECK(p.compensations == nullptr, "rs_render_frame: antialiased mode is not available on the fused path");
    RS_CHECK(a->max_isects > 0 && a->max_isects < ((int64_t)1 << 31), "rs_render_frame: bad max_isects");
    p.tile_width = (p.image_width + p.tile_size - 1) / p.tile_size;
    p.tile_height = (p.image_height + p.tile_size - 1) / p.tile_size;
    const FrameLayout L = make_layout(p.C, p.N, p.image_width, p.image_height, p.tile_size, a->max_isects);
    RS_CHECK(a->workspace_bytes >= L.total, "rs_render_frame: workspace too small (%llu < %llu)",
             (unsigned long long)a->workspace_bytes, (unsigned long long)L.total);
    char *w = reinterpret_cast<char *>(a->workspace);
    cudaStream_t s = (cudaStream_t)stream;

    p.radii = reinterpret_cast<int32_t *>(w + L.radii);
    p.means2d = reinterpret_cast<float *>(w + L.means2d);
    p.depths = reinterpret_cast<float *>(w + L.depths);
    p.conics = reinterpret_cast<float *>(w + L.conics);
    p.records = reinterpret_cast<float *>(w + L.records);
    if (p.sh_coeffs != nullptr)
        p.sh_colors = reinterpret_cast<float *>(w + L.sh_colors);
    p.tiles_per_gauss = reinterpret_cast<int32_t *>(w + L.tiles_per_gauss);
    p.block_sums = nullptr; // the depth-ordered binning computes its own block sums
    p.tile_footprints = a->tight_tiles != 0 ? reinterpret_cast<uint32_t *>(w + L.footprints) : nullptr;
    RS_CHECK(a->stages >= 0 && a->stages <= 3, "rs_render_frame: bad stages %d", a->stages);
    const bool do_bin = a->stages == 0 || (a->stages & RS_FRAME_BIN), do_composite = a->stages == 0 || (a->stages & RS_FRAME_COMPOSITE);
    if (ev)
        RS_CUDA(cudaEventRecord(ev[0], s));
    // the depth statistics the ordering starts from (min / max of the visible depth bits, their number) come out of the
    // projection kernel: one kernel and one pass over depths + tile counts less per frame
    const int64_t E = (int64_t)p.C * p.N;
    if (do_bin && E > 0) {
        if (int e = rs_isect_sorted_prepare(w + L.bin_ws, E, a->max_isects, stream))
            return e;
        p.depth_stats = rs_isect_sorted_depth_stats(w + L.bin_ws, E, a->max_isects);
    }
    if (do_bin)
        if (int e = rs_project_fwd(&p, stream))
            return e;
    if (ev)
        RS_CUDA(cudaEventRecord(ev[1], s));

    rs_isect_sorted_args sa;
    memset(&sa, 0, sizeof(sa));
    fill_isect(sa.isect, p, a, w, L);
    int32_t *offsets = a->out_tile_offsets != nullptr ? a->out_tile_offsets : reinterpret_cast<int32_t *>(w + L.tile_offsets);
    sa.tile_offsets = offsets;
    sa.workspace = w + L.bin_ws;
    sa.workspace_bytes = L.bin_ws_bytes;
    sa.depth_stats_ready = p.depth_stats != nullptr ? 1 : 0;
    sa.tile_footprints = p.tile_footprints;
    if (do_bin)
        if (int e = rs_isect_sorted(&sa, stream))
            return e;
    const int32_t *vals_sorted = sa.isect.flatten_ids;
    if (ev)
        RS_CUDA(cudaEventRecord(ev[2], s));

    rs_raster_fwd_args r;
    memset(&r, 0, sizeof(r));
    r.I = p.C;
    r.N = p.N;
    r.channels = a->channels;
    r.image_width = p.image_width;
    r.image_height = p.image_height;
    r.tile_size = p.tile_size;
    r.tile_width = p.tile_width;
    r.tile_height = p.tile_height;
    r.n_isects = a->max_isects;
    r.n_isects_dev = a->status;
    r.means2d = p.means2d;
    r.conics = p.conics;
    r.colors = p.sh_coeffs != nullptr ? p.sh_colors : a->colors;
    r.opacities = p.opacities;
    r.attr_mod_colors = (a->colors_per_camera || p.sh_coeffs != nullptr) ? 0 : p.N;
    r.attr_mod_opacities = p.N;
    r.backgrounds = a->backgrounds;
    r.masks = nullptr;
    r.tile_offsets = offsets;
    r.flatten_ids = vals_sorted;
    r.render_colors = a->render_colors;
    r.render_alphas = a->render_alphas;
    r.render_rgb8 = a->render_rgb8;
    r.last_ids = reinterpret_cast<int32_t *>(w + L.last_ids);
    r.records = p.records;
    r.records_ready = 1;
    r.n_rows = (int64_t)p.C * p.N;
    r.tile_counter = reinterpret_cast<uint32_t *>(w + L.tile_counter);
    if (do_composite)
        RS_CUDA(cudaMemsetAsync(w + L.tile_counter, 0, 4, s));
    if (do_composite)
        if (int e = rs_raster_fwd(&r, stream))
            return e;
    if (ev)
        RS_CUDA(cudaEventRecord(ev[3], s));
    return 0;
}
