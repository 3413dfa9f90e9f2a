// rs_seghead_fwd / rs_seghead_bwd: the segmentation head of the identity-feature training step, fused.
// Replaces `torch.nn.Sequential(Linear(D, 64), ReLU(), Linear(64, D))` applied to the per-Gaussian identity encodings
// (examples/simple_trainer.py:442-446, 946-947: processed = segmentation_head(raw_identities), N = 1 M rows) and its autograd.
// In torch that is two library GEMMs + an activation forward and four GEMMs + masks backward, with the [N, 64] hidden tensor
// written to and re-read from HBM five times (256 MB a pass at 1 M Gaussians).  Here the hidden layer only ever exists in
// registers / shared memory:
//   forward   one thread per Gaussian, weights in shared memory (broadcast reads), 2 * D * H FMAs per row;
//             HBM traffic = x in, y out (2 * N * D * 4 B).
//   backward  a CTA takes 128 rows at a time: every thread recomputes its row's hidden activations, forms v_h = relu'(a) *
//             W2^T v_y and v_x = W1^T v_h, and parks h / v_h / x / v_y of the chunk in shared memory (column-major with an odd
//             pitch: conflict-free both ways); then the weight gradients v_W2 += v_y^T h, v_W1 += v_h^T x (two D x H tile
//             products over the 128 rows) are accumulated in registers -- a 2 x 8 tile of v_W2 in each of 64 threads, a 4 x 4 tile
//             of v_W1 in each of the other 64 -- across all chunks of a persistent CTA, and flushed with one atomicAdd per entry
//             and CTA.
// float32 throughout, FMA order fixed per row (deterministic forward; weight gradients differ only by the atomic flush order).
#include "common.cuh"

#define SEG_THREADS 128
#define SEG_MAX_D 32
#define SEG_MAX_H 128

struct SegWeights { // shared-memory image of the four parameter tensors; w2 is held TRANSPOSED ([H][D]) so that the
    float *w1, *b1, *w2, *b2; // weights of one hidden unit are one contiguous row in both matrices (float4 broadcasts)
};
__device__ __forceinline__ SegWeights seg_load_weights(const rs_seghead_args &a, float *smem) {
    SegWeights w;
    const int D = a.D, H = a.H;
    w.w1 = smem;
    w.b1 = w.w1 + H * D;
    w.w2 = w.b1 + H;
    w.b2 = w.w2 + D * H;
    for (int i = threadIdx.x; i < H * D; i += blockDim.x) {
        w.w1[i] = a.w1[i];
        const int o = i / H, j = i - o * H; // a.w2 is [D][H]
        w.w2[j * D + o] = a.w2[i];
    }
    for (int i = threadIdx.x; i < H; i += blockDim.x)
        w.b1[i] = a.b1[i];
    for (int i = threadIdx.x; i < D; i += blockDim.x)
        w.b2[i] = a.b2[i];
    return w;
}
__host__ __device__ static inline size_t seg_weight_floats(int D, int H) { return (size_t)2 * H * D + H + D; }

template <int D, int H>
__global__ void __launch_bounds__(SEG_THREADS) rs_seghead_fwd_kernel(const rs_seghead_args a) {
    extern __shared__ __align__(16) float seg_smem[];
    const SegWeights w = seg_load_weights(a, seg_smem);
    __syncthreads();
    for (int64_t n = (int64_t)blockIdx.x * SEG_THREADS + threadIdx.x; n < a.N; n += (int64_t)gridDim.x * SEG_THREADS) {
        float x[D], y[D];
#pragma unroll
        for (int i = 0; i < D; i += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(a.x + n * D + i);
            x[i] = v.x, x[i + 1] = v.y, x[i + 2] = v.z, x[i + 3] = v.w;
        }
#pragma unroll
        for (int o = 0; o < D; ++o)
            y[o] = w.b2[o];
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            float h = w.b1[j];
            const float4 *r1 = reinterpret_cast<const float4 *>(w.w1 + j * D);
            const float4 *r2 = reinterpret_cast<const float4 *>(w.w2 + j * D);
#pragma unroll
            for (int i = 0; i < D; i += 4) {
                const float4 q = r1[i >> 2];
                h = fmaf(q.x, x[i], h);
                h = fmaf(q.y, x[i + 1], h);
                h = fmaf(q.z, x[i + 2], h);
                h = fmaf(q.w, x[i + 3], h);
            }
            h = fmaxf(h, 0.f);
#pragma unroll
            for (int o = 0; o < D; o += 4) {
                const float4 q = r2[o >> 2];
                y[o] = fmaf(q.x, h, y[o]);
                y[o + 1] = fmaf(q.y, h, y[o + 1]);
                y[o + 2] = fmaf(q.z, h, y[o + 2]);
                y[o + 3] = fmaf(q.w, h, y[o + 3]);
            }
        }
#pragma unroll
        for (int o = 0; o < D; o += 4)
            *reinterpret_cast<float4 *>(a.y + n * D + o) = make_float4(y[o], y[o + 1], y[o + 2], y[o + 3]);
    }
}

// generic sizes (D <= 32, H <= 128, D % 4 == 0 not required): same arithmetic with run-time loops
__global__ void __launch_bounds__(SEG_THREADS) rs_seghead_fwd_generic_kernel(const rs_seghead_args a) {
    extern __shared__ __align__(16) float seg_smem[];
    const SegWeights w = seg_load_weights(a, seg_smem);
    __syncthreads();
    const int D = a.D, H = a.H;
    for (int64_t n = (int64_t)blockIdx.x * SEG_THREADS + threadIdx.x; n < a.N; n += (int64_t)gridDim.x * SEG_THREADS) {
        float x[SEG_MAX_D], y[SEG_MAX_D];
        for (int i = 0; i < D; ++i) {
            x[i] = a.x[n * D + i];
            y[i] = w.b2[i];
        }
        for (int j = 0; j < H; ++j) {
            float h = w.b1[j];
            for (int i = 0; i < D; ++i)
                h = fmaf(w.w1[j * D + i], x[i], h);
            h = fmaxf(h, 0.f);
            for (int o = 0; o < D; ++o)
                y[o] = fmaf(w.w2[j * D + o], h, y[o]);
        }
        for (int o = 0; o < D; ++o)
            a.y[n * D + o] = y[o];
    }
}

// ---- backward, D = 16, H = 64 ---------------------------------------------------------------------------------------------
#define SEG_CHUNK SEG_THREADS
#define SEG_PITCH (SEG_CHUNK + 1) // odd pitch: column-major tiles are conflict-free for row-wise writes and column-wise reads
template <int D, int H>
__global__ void __launch_bounds__(SEG_THREADS) rs_seghead_bwd_kernel(const rs_seghead_args a) {
    static_assert(D == 16 && H == 64 && SEG_THREADS == 128, "thread -> gradient-entry mapping below assumes 16 x 64 x 128");
    extern __shared__ __align__(16) float seg_smem[];
    const SegWeights w = seg_load_weights(a, seg_smem);
    float *hs = seg_smem + ((seg_weight_floats(D, H) + 3) & ~(size_t)3); // [H][PITCH]
    float *vhs = hs + H * SEG_PITCH;                                      // [H][PITCH]
    float *xs = vhs + H * SEG_PITCH;                                      // [D][PITCH]
    float *vys = xs + D * SEG_PITCH;                                      // [D][PITCH]
    __syncthreads();
    const int t = threadIdx.x;
    // gradient entries owned by this thread, a 2 x 8 (threads 0..63: v_W2[oa .. oa+2)[ja .. ja+8)) or 4 x 4 (threads 64..127:
    // v_W1[ja .. ja+4)[oa .. oa+4)) register tile: 10 / 8 shared-memory loads per 16 FMAs and row of the chunk
    const bool on_w2 = t < 64;
    const int u = t & 63;
    const int oa = on_w2 ? (u >> 3) * 2 : (u & 3) * 4; // v_W2: output row pair      | v_W1: input column quad
    const int ja = on_w2 ? (u & 7) * 8 : (u >> 2) * 4; // v_W2: hidden column octet  | v_W1: hidden row quad
    float g[16], gb[4];
#pragma unroll
    for (int q = 0; q < 16; ++q)
        g[q] = 0.f;
    gb[0] = gb[1] = gb[2] = gb[3] = 0.f;

    const int64_t n_chunks = (a.N + SEG_CHUNK - 1) / SEG_CHUNK;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t n = c * SEG_CHUNK + t;
        const bool live = n < a.N;
        float x[D], vy[D];
#pragma unroll
        for (int i = 0; i < D; i += 4) {
            const float4 v = live ? *reinterpret_cast<const float4 *>(a.x + n * D + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 u = live ? *reinterpret_cast<const float4 *>(a.v_y + n * D + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            x[i] = v.x, x[i + 1] = v.y, x[i + 2] = v.z, x[i + 3] = v.w;
            vy[i] = u.x, vy[i + 1] = u.y, vy[i + 2] = u.z, vy[i + 3] = u.w;
        }
        float vx[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            vx[i] = 0.f;
            xs[i * SEG_PITCH + t] = x[i];
            vys[i * SEG_PITCH + t] = vy[i];
        }
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            float pre = w.b1[j], vh = 0.f;
            float w1r[D];
            const float4 *r1 = reinterpret_cast<const float4 *>(w.w1 + j * D);
            const float4 *r2 = reinterpret_cast<const float4 *>(w.w2 + j * D);
#pragma unroll
            for (int i = 0; i < D; i += 4) {
                const float4 q = r1[i >> 2];
                w1r[i] = q.x, w1r[i + 1] = q.y, w1r[i + 2] = q.z, w1r[i + 3] = q.w;
                const float4 u = r2[i >> 2];
                vh = fmaf(u.x, vy[i], vh);
                vh = fmaf(u.y, vy[i + 1], vh);
                vh = fmaf(u.z, vy[i + 2], vh);
                vh = fmaf(u.w, vy[i + 3], vh);
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
                pre = fmaf(w1r[i], x[i], pre);
            vh = (pre > 0.f && live) ? vh : 0.f;
            hs[j * SEG_PITCH + t] = live ? fmaxf(pre, 0.f) : 0.f;
            vhs[j * SEG_PITCH + t] = vh;
#pragma unroll
            for (int i = 0; i < D; ++i)
                vx[i] = fmaf(w1r[i], vh, vx[i]);
        }
        if (live && a.v_x != nullptr) {
#pragma unroll
            for (int i = 0; i < D; i += 4)
                *reinterpret_cast<float4 *>(a.v_x + n * D + i) = make_float4(vx[i], vx[i + 1], vx[i + 2], vx[i + 3]);
        }
        __syncthreads();
        // weight-gradient tile products over the rows of the chunk
        if (on_w2) {
            const float *pa = vys + oa * SEG_PITCH, *pb = hs + ja * SEG_PITCH;
#pragma unroll 4
            for (int r = 0; r < SEG_CHUNK; ++r) {
                const float a0 = pa[r], a1 = pa[SEG_PITCH + r];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float hq = pb[q * SEG_PITCH + r];
                    g[q] = fmaf(a0, hq, g[q]);
                    g[8 + q] = fmaf(a1, hq, g[8 + q]);
                }
                if (ja == 0) { // one thread per output-row pair also sums the bias gradient
                    gb[0] += a0;
                    gb[1] += a1;
                }
            }
        } else {
            const float *pa = vhs + ja * SEG_PITCH, *pb = xs + oa * SEG_PITCH;
#pragma unroll 4
            for (int r = 0; r < SEG_CHUNK; ++r) {
                float av[4], bv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    av[q] = pa[q * SEG_PITCH + r];
                    bv[q] = pb[q * SEG_PITCH + r];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        g[q * 4 + p] = fmaf(av[q], bv[p], g[q * 4 + p]);
                if (oa == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        gb[q] += av[q];
                }
            }
        }
        __syncthreads();
    }
    if (on_w2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            atomicAdd(a.v_w2 + oa * H + ja + q, g[q]);
            atomicAdd(a.v_w2 + (oa + 1) * H + ja + q, g[8 + q]);
        }
        if (ja == 0) {
            atomicAdd(a.v_b2 + oa, gb[0]);
            atomicAdd(a.v_b2 + oa + 1, gb[1]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
                atomicAdd(a.v_w1 + (ja + q) * D + oa + p, g[q * 4 + p]);
            if (oa == 0)
                atomicAdd(a.v_b1 + ja + q, gb[q]);
        }
    }
}

static int seg_check(const rs_seghead_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->N >= 0 && a->D >= 1 && a->D <= SEG_MAX_D && a->H >= 1 && a->H <= SEG_MAX_H, "%s: bad sizes (N %lld, D %d, H %d)",
             who, (long long)a->N, a->D, a->H);
    RS_CHECK(a->w1 && a->b1 && a->w2 && a->b2, "%s: null parameter pointer", who);
    return 0;
}

extern "C" int rs_seghead_fwd(const rs_seghead_args *a, rs_stream_t stream) {
    if (int e = seg_check(a, "rs_seghead_fwd"))
        return e;
    if (a->N == 0)
        return 0;
    RS_CHECK(a->x && a->y, "rs_seghead_fwd: null x / y");
    const size_t smem = seg_weight_floats(a->D, a->H) * sizeof(float);
    const int grid = (int)min((int64_t)rs_num_sms() * 8, (a->N + SEG_THREADS - 1) / SEG_THREADS);
    const bool aligned = ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->y)) & 15) == 0;
    if (a->D == 16 && a->H == 64 && aligned)
        rs_seghead_fwd_kernel<16, 64><<<grid, SEG_THREADS, smem, (cudaStream_t)stream>>>(*a);
    else
        rs_seghead_fwd_generic_kernel<<<grid, SEG_THREADS, smem, (cudaStream_t)stream>>>(*a);
    RS_LAUNCH_CHECK("rs_seghead_fwd_kernel");
    return 0;
}

// v_w1 / v_b1 / v_w2 / v_b2 are ACCUMULATED into (zero-initialised by the caller); v_x is written (optional)
extern "C" int rs_seghead_bwd(const rs_seghead_args *a, rs_stream_t stream) {
    if (int e = seg_check(a, "rs_seghead_bwd"))
        return e;
    RS_CHECK(a->D == 16 && a->H == 64, "rs_seghead_bwd: the fused backward is built for identity_dim 16 and 64 hidden units "
                                       "(examples/simple_trainer.py:442-446), got D %d, H %d", a->D, a->H);
    if (a->N == 0)
        return 0;
    RS_CHECK(a->x && a->v_y && a->v_w1 && a->v_b1 && a->v_w2 && a->v_b2, "rs_seghead_bwd: null pointer");
    RS_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->v_y) | reinterpret_cast<uintptr_t>(a->v_x)) & 15) == 0,
             "rs_seghead_bwd: x / v_y / v_x must be 16-byte aligned");
    const size_t floats = ((seg_weight_floats(16, 64) + 3) & ~(size_t)3) + (size_t)2 * 64 * SEG_PITCH + (size_t)2 * 16 * SEG_PITCH;
    const size_t smem = floats * sizeof(float);
    static RsPerDevice attr;
    if (!rs_dev_done(attr)) {
        RS_CUDA(cudaFuncSetAttribute(rs_seghead_bwd_kernel<16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rs_dev_mark(attr);
    }
    const int64_t n_chunks = (a->N + SEG_CHUNK - 1) / SEG_CHUNK;
    const int grid = (int)min((int64_t)rs_num_sms() * 2, n_chunks);
    rs_seghead_bwd_kernel<16, 64><<<grid, SEG_THREADS, smem, (cudaStream_t)stream>>>(*a);
    RS_LAUNCH_CHECK("rs_seghead_bwd_kernel");
    return 0;
}
