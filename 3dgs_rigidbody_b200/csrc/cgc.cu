// rs_cgc_*: contrastive clustering loss of the identity-feature training step, forward and backward.
// Replaces `cgc_contrastive_clustering_loss` (examples/utils.py:828-904; called from examples/simple_trainer.py:945-975 on
// the [H, W, 16] feature map that the c3 compositing pass renders), which is ~25 torch kernels forward and as many again
// in autograd, each a pass over the 133 MB feature map.  Here the feature map is read three times forward and once more
// backward; everything cluster-sized (centroids, temperatures, their gradients: at most RS_CGC_MAX_CLUSTERS x
// RS_CGC_MAX_DIM floats) is recomputed in the prologue of the kernel that needs it, in shared memory.
//
// With f_p = x_p / max(|x_p|, 1e-12), members M_k of cluster k (n_k pixels), c_k = normalize(mean_{p in M_k} f_p),
// active pixels p with target t_p (A of them, a_k per target), s_pk = f_p . c_k, phi_k = max(mean_{t_p = k} s_pk, eps):
//     loss = 1/A sum_p [ logsumexp_k(s_pk / phi_{t_p}) - s_{p,t_p} / phi_{t_p} ]
// `target` and `member` differ only through a quirk of the reference that is reproduced on purpose: background pixels
// index its cluster table with -1, i.e. they become ACTIVE pixels of the LAST foreground cluster (utils.py:878-883) without
// being members of it (they are not in the centroid, utils.py:859-862).  The host side derives both arrays from the mask.
//
// Backward (derivation checked against the reference's autograd, tests/golden/make_golden_cgc.py):
//     g_pk = (softmax_k(s_pk / tau_p) - [k = t_p]) / (tau_p A),   tau_p = phi_{t_p}
//     h_k  = [phi_k > eps] sum_{t_p = k} ( - sum_j g_pj s_pj / tau_p )
//     G_pk = g_pk + [k = t_p] h_k / a_k
//     u_k  = sum_p G_pk f_p = U0_k + (h_k / a_k) Sact_k,   v_k = (u_k - (u_k . c_k) c_k) / (|m_k| n_k)
//     dL/df_p = sum_k G_pk c_k + [p in M_k] v_k,           dL/dx_p = (dL/df_p - (dL/df_p . f_p) f_p) / |x_p|
#include "common.cuh"

#define CGC_THREADS 128
#define CGC_TILE 128 // pixels per CTA iteration (one per thread)

// Per-pixel feature vectors live in registers: every loop over channels has the compile-time bound DM (16 or 32) and is
// fully unrolled; channels beyond the runtime D hold zeros, so no guard is needed inside dot products.
template <int DM> __device__ __forceinline__ float cgc_normalize(const float (&x)[DM], float (&f)[DM]) {
    float n2 = 0.f;
#pragma unroll
    for (int d = 0; d < DM; ++d)
        n2 = fmaf(x[d], x[d], n2);
    const float nrm = fmaxf(sqrtf(n2), 1e-12f); // F.normalize: x / max(|x|, eps)
    const float inv = 1.f / nrm;
#pragma unroll
    for (int d = 0; d < DM; ++d)
        f[d] = x[d] * inv;
    return nrm;
}

// f . c_k for a centroid row in shared memory (row pitch D)
template <int DM> __device__ __forceinline__ float cgc_dot(const float (&f)[DM], const float *row, int D) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DM; ++d)
        if (d < D)
            s = fmaf(f[d], row[d], s);
    return s;
}

// workspace layout (floats): [S_member K*D | S_active K*D | U0 K*D | possum K | h K | loss 1]
__device__ __forceinline__ float *ws_smem(const rs_cgc_args &a) { return a.ws; }
__device__ __forceinline__ float *ws_sact(const rs_cgc_args &a) { return a.ws + (size_t)a.K * a.D; }
__device__ __forceinline__ float *ws_u0(const rs_cgc_args &a) { return a.ws + 2 * (size_t)a.K * a.D; }
__device__ __forceinline__ float *ws_possum(const rs_cgc_args &a) { return a.ws + 3 * (size_t)a.K * a.D; }
__device__ __forceinline__ float *ws_h(const rs_cgc_args &a) { return ws_possum(a) + a.K; }
__device__ __forceinline__ float *ws_loss(const rs_cgc_args &a) { return ws_h(a) + a.K; }

// centroids from the member sums: c_k = normalize(S_k / n_k)
__device__ __forceinline__ void cgc_centroids(const rs_cgc_args &a, float *c, float *mn) {
    const int K = a.K, D = a.D;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float inv_n = 1.f / a.n_member[k];
        float n2 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float m = ws_smem(a)[k * D + d] * inv_n;
            n2 = fmaf(m, m, n2);
        }
        const float nrm = fmaxf(sqrtf(n2), 1e-12f);
        mn[k] = nrm;
        for (int d = 0; d < D; ++d)
            c[k * D + d] = ws_smem(a)[k * D + d] * inv_n / nrm;
    }
}

__device__ __forceinline__ void cgc_phi(const rs_cgc_args &a, float *phi) {
    for (int k = threadIdx.x; k < a.K; k += blockDim.x)
        phi[k] = fmaxf(ws_possum(a)[k] / fmaxf(a.n_active[k], 1.f), a.eps);
}

template <int DM> __device__ __forceinline__ void cgc_load_pixel(const rs_cgc_args &a, int64_t p, float (&x)[DM]) {
    const float *src = a.features + p * a.D;
    if ((a.D & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.features) & 15) == 0)) {
#pragma unroll
        for (int d = 0; d < DM; d += 4) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d < a.D)
                q = *reinterpret_cast<const float4 *>(src + d);
            x[d] = q.x, x[d + 1] = q.y, x[d + 2] = q.z, x[d + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < DM; ++d)
            x[d] = d < a.D ? src[d] : 0.f;
    }
}

// ---- pass 1: per-cluster sums of the normalised features (members and active pixels) ---------------------------------
template <int DM> __global__ void __launch_bounds__(CGC_THREADS) rs_cgc_sums_kernel(const rs_cgc_args a) {
    extern __shared__ float sm[];
    const int K = a.K, D = a.D;
    float *s_mem = sm, *s_act = sm + K * D;
    for (int i = threadIdx.x; i < 2 * K * D; i += blockDim.x)
        sm[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < a.P; p0 += stride) { // warp-uniform trip count
        const int64_t p = p0 + threadIdx.x;
        const int t = p < a.P ? a.target[p] : -1, m = p < a.P ? a.member[p] : -1;
        float x[DM], f[DM];
#pragma unroll
        for (int d = 0; d < DM; ++d)
            f[d] = 0.f;
        if (t >= 0 || m >= 0) {
            cgc_load_pixel<DM>(a, p, x);
            cgc_normalize<DM>(x, f);
        }
        // instance masks are spatially coherent: most warps sit inside ONE cluster, and then 32 lanes would hammer the same
        // D shared-memory addresses.  Uniform warps reduce with shuffles and issue one atomic per channel instead.
        const int t0 = __shfl_sync(0xffffffffu, t, 0), m0 = __shfl_sync(0xffffffffu, m, 0);
        if (__all_sync(0xffffffffu, t == t0 && m == m0)) {
            if (t0 < 0 && m0 < 0)
                continue;
#pragma unroll
            for (int d = 0; d < DM; ++d) {
                float v = f[d];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                f[d] = v;
            }
            if (lane == 0) {
#pragma unroll
                for (int d = 0; d < DM; ++d)
                    if (d < D) {
                        if (m0 >= 0)
                            atomicAdd(&s_mem[m0 * D + d], f[d]);
                        if (t0 >= 0)
                            atomicAdd(&s_act[t0 * D + d], f[d]);
                    }
            }
        } else {
#pragma unroll
            for (int d = 0; d < DM; ++d)
                if (d < D) {
                    if (m >= 0)
                        atomicAdd(&s_mem[m * D + d], f[d]);
                    if (t >= 0)
                        atomicAdd(&s_act[t * D + d], f[d]);
                }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * K * D; i += blockDim.x)
        if (sm[i] != 0.f)
            atomicAdd(&a.ws[i], sm[i]);
}

// ---- pass 2: sum of the positive similarities per target (-> temperatures) ---------------------------------------------
template <int DM> __global__ void __launch_bounds__(CGC_THREADS) rs_cgc_pos_kernel(const rs_cgc_args a) {
    extern __shared__ float sm[];
    const int K = a.K, D = a.D;
    float *c = sm, *mn = c + K * D, *ps = mn + K;
    cgc_centroids(a, c, mn);
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        ps[k] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < a.P; p0 += stride) {
        const int64_t p = p0 + threadIdx.x;
        const int t = p < a.P ? a.target[p] : -1;
        float s = 0.f;
        if (t >= 0) {
            float x[DM], f[DM];
            cgc_load_pixel<DM>(a, p, x);
            cgc_normalize<DM>(x, f);
            s = cgc_dot<DM>(f, c + t * D, D);
        }
        const int t0 = __shfl_sync(0xffffffffu, t, 0);
        if (__all_sync(0xffffffffu, t == t0)) { // one cluster per warp: a shuffle reduction and a single atomic
            if (t0 < 0)
                continue;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0)
                atomicAdd(&ps[t0], s);
        } else if (t >= 0) {
            atomicAdd(&ps[t], s);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (ps[k] != 0.f)
            atomicAdd(&ws_possum(a)[k], ps[k]);
}

// ---- pass 3: the loss; with accumulate_grad also U0 = sum_p g_p (x) f_p and h ------------------------------------------
template <int DM> __global__ void __launch_bounds__(CGC_THREADS) rs_cgc_loss_kernel(const rs_cgc_args a) {
    extern __shared__ float sm[];
    const int K = a.K, D = a.D;
    float *c = sm, *mn = c + K * D, *phi = mn + K, *hs = phi + K;
    float *ft = hs + K;                       // [CGC_TILE][D + 1]   (backward only)
    float *gt = ft + CGC_TILE * (D + 1);      // [CGC_TILE][K + 1]   (backward only)
    __shared__ float red[CGC_THREADS / 32];
    cgc_centroids(a, c, mn);
    cgc_phi(a, phi);
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        hs[k] = 0.f;
    __syncthreads();
    const bool bwd = a.accumulate_grad != 0;
    const float inv_A = 1.f / (float)a.A;
    // outputs of the U0 tile product handled by this thread: o = threadIdx.x + i * CGC_THREADS -> (k, d) = (o / D, o % D)
    float acc[RS_CGC_MAX_CLUSTERS * RS_CGC_MAX_DIM / CGC_THREADS];
#pragma unroll
    for (int i = 0; i < RS_CGC_MAX_CLUSTERS * RS_CGC_MAX_DIM / CGC_THREADS; ++i)
        acc[i] = 0.f;
    float loss_local = 0.f;
    const int64_t n_tiles = (a.P + CGC_TILE - 1) / CGC_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p = tile * CGC_TILE + threadIdx.x;
        const int t = p < a.P ? a.target[p] : -1;
        float f[DM];
        if (t >= 0) {
            float x[DM];
            cgc_load_pixel<DM>(a, p, x);
            cgc_normalize<DM>(x, f);
            const float inv_tau = 1.f / phi[t];
            // logits and a max-subtracted logsumexp
            float mx = -3e38f;
            for (int k = 0; k < K; ++k) {
                const float s = cgc_dot<DM>(f, c + k * D, D);
                if (bwd)
                    gt[threadIdx.x * (K + 1) + k] = s; // similarities, turned into g below
                mx = fmaxf(mx, s * inv_tau);
            }
            float den = 0.f, st = 0.f;
            for (int k = 0; k < K; ++k) {
                const float s = bwd ? gt[threadIdx.x * (K + 1) + k] : cgc_dot<DM>(f, c + k * D, D);
                den += expf(s * inv_tau - mx);
                if (k == t)
                    st = s;
            }
            loss_local += (logf(den) + mx - st * inv_tau) * inv_A;
            if (bwd) {
                float gs = 0.f;
                const float inv_den = 1.f / den;
                for (int k = 0; k < K; ++k) {
                    const float s = gt[threadIdx.x * (K + 1) + k];
                    const float q = expf(s * inv_tau - mx) * inv_den;
                    const float g = (q - (k == t ? 1.f : 0.f)) * inv_tau * inv_A;
                    gt[threadIdx.x * (K + 1) + k] = g;
                    gs = fmaf(g, s, gs);
                }
                atomicAdd(&hs[t], -gs * inv_tau);
#pragma unroll
                for (int d = 0; d < DM; ++d)
                    if (d < D)
                        ft[threadIdx.x * (D + 1) + d] = f[d];
            }
        } else if (bwd) {
            for (int k = 0; k < K; ++k)
                gt[threadIdx.x * (K + 1) + k] = 0.f;
            for (int d = 0; d < D; ++d)
                ft[threadIdx.x * (D + 1) + d] = 0.f;
        }
        if (bwd) { // U0 += G_tile^T F_tile, accumulated in registers across all tiles of this CTA
            __syncthreads();
#pragma unroll
            for (int i = 0; i < RS_CGC_MAX_CLUSTERS * RS_CGC_MAX_DIM / CGC_THREADS; ++i) {
                const int o = threadIdx.x + i * CGC_THREADS;
                if (o < K * D) {
                    const int k = o / D, d = o - k * D;
                    float s = acc[i];
                    for (int q = 0; q < CGC_TILE; ++q)
                        s = fmaf(gt[q * (K + 1) + k], ft[q * (D + 1) + d], s);
                    acc[i] = s;
                }
            }
            __syncthreads();
        }
    }
    // loss: warp shuffle, then one atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = loss_local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < CGC_THREADS / 32; ++w)
            s += red[w];
        atomicAdd(ws_loss(a), s);
    }
    if (bwd) {
#pragma unroll
        for (int i = 0; i < RS_CGC_MAX_CLUSTERS * RS_CGC_MAX_DIM / CGC_THREADS; ++i) {
            const int o = threadIdx.x + i * CGC_THREADS;
            if (o < K * D && acc[i] != 0.f)
                atomicAdd(&ws_u0(a)[o], acc[i]);
        }
        for (int k = threadIdx.x; k < K; k += blockDim.x)
            if (hs[k] != 0.f)
                atomicAdd(&ws_h(a)[k], hs[k]);
    }
}

// ---- pass 4: dL/dx per pixel ---------------------------------------------------------------------------------------------
template <int DM> __global__ void __launch_bounds__(CGC_THREADS) rs_cgc_grad_kernel(const rs_cgc_args a) {
    extern __shared__ float sm[];
    const int K = a.K, D = a.D;
    float *c = sm, *mn = c + K * D, *phi = mn + K, *hk = phi + K, *v = hk + K;
    cgc_centroids(a, c, mn);
    cgc_phi(a, phi);
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float na = fmaxf(a.n_active[k], 1.f);
        const float phi_raw = ws_possum(a)[k] / na;
        hk[k] = phi_raw > a.eps ? ws_h(a)[k] / na : 0.f; // clamp_min passes the gradient only above the floor
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) { // v_k = (u_k - (u_k . c_k) c_k) / (|m_k| n_k)
        float dot = 0.f;
        for (int d = 0; d < D; ++d) {
            const float u = ws_u0(a)[k * D + d] + hk[k] * ws_sact(a)[k * D + d];
            dot = fmaf(u, c[k * D + d], dot);
        }
        const float scale = 1.f / (mn[k] * a.n_member[k]);
        for (int d = 0; d < D; ++d) {
            const float u = ws_u0(a)[k * D + d] + hk[k] * ws_sact(a)[k * D + d];
            v[k * D + d] = (u - dot * c[k * D + d]) * scale;
        }
    }
    __syncthreads();
    const float inv_A = 1.f / (float)a.A;
    const float up = a.grad_loss != nullptr ? *a.grad_loss : 1.f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += (int64_t)gridDim.x * blockDim.x) {
        const int t = a.target[p], m = a.member[p];
        float *out = a.v_features + p * D;
        if (t < 0 && m < 0) {
            for (int d = 0; d < D; ++d)
                out[d] = 0.f;
            continue;
        }
        float x[DM], f[DM], df[DM];
        cgc_load_pixel<DM>(a, p, x);
        const float nrm = cgc_normalize<DM>(x, f);
#pragma unroll
        for (int d = 0; d < DM; ++d)
            df[d] = (m >= 0 && d < D) ? v[m * D + d] : 0.f;
        if (t >= 0) {
            const float inv_tau = 1.f / phi[t];
            float mx = -3e38f;
            for (int k = 0; k < K; ++k)
                mx = fmaxf(mx, cgc_dot<DM>(f, c + k * D, D) * inv_tau);
            float den = 0.f;
            for (int k = 0; k < K; ++k)
                den += expf(cgc_dot<DM>(f, c + k * D, D) * inv_tau - mx);
            const float inv_den = 1.f / den;
            for (int k = 0; k < K; ++k) {
                const float q = expf(cgc_dot<DM>(f, c + k * D, D) * inv_tau - mx) * inv_den;
                float G = (q - (k == t ? 1.f : 0.f)) * inv_tau * inv_A;
                if (k == t)
                    G += hk[k];
#pragma unroll
                for (int d = 0; d < DM; ++d)
                    if (d < D)
                        df[d] = fmaf(G, c[k * D + d], df[d]);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < DM; ++d)
            dot = fmaf(df[d], f[d], dot);
        const float inv = up / nrm;
        const bool tiny = nrm <= 1e-12f; // below F.normalize's floor the norm is a constant
#pragma unroll
        for (int d = 0; d < DM; ++d)
            if (d < D)
                out[d] = (tiny ? df[d] : df[d] - dot * f[d]) * inv;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
extern "C" uint64_t rs_cgc_workspace_floats(int32_t K, int32_t D) { return 3ull * K * D + 2ull * K + 1; }

static int cgc_check(const rs_cgc_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->P >= 0 && a->P < ((int64_t)1 << 40), "%s: bad pixel count", who);
    RS_CHECK(a->D >= 1 && a->D <= RS_CGC_MAX_DIM, "%s: feature dim %d outside 1..%d", who, a->D, RS_CGC_MAX_DIM);
    RS_CHECK(a->K >= 2 && a->K <= RS_CGC_MAX_CLUSTERS, "%s: %d clusters outside 2..%d", who, a->K, RS_CGC_MAX_CLUSTERS);
    RS_CHECK(a->A >= 1, "%s: no active pixel", who);
    RS_CHECK(a->features && a->target && a->member && a->n_member && a->n_active && a->ws, "%s: null pointer", who);
    return 0;
}

static int cgc_grid(int64_t units) {
    const int64_t want = (units + CGC_THREADS - 1) / CGC_THREADS;
    const int64_t cap = (int64_t)rs_num_sms() * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <typename Kern> static int cgc_launch(Kern kern, const rs_cgc_args *a, size_t smem, int grid, cudaStream_t s, const char *name) {
    if (smem > 48 * 1024)
        RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, CGC_THREADS, smem, s>>>(*a);
    RS_LAUNCH_CHECK(name);
    return 0;
}

template <int DM> static int cgc_fwd_impl(const rs_cgc_args *a, cudaStream_t s) {
    const size_t KD = (size_t)a->K * a->D;
    if (int e = cgc_launch(rs_cgc_sums_kernel<DM>, a, 2 * KD * sizeof(float), cgc_grid(a->P), s, "rs_cgc_sums_kernel"))
        return e;
    if (int e = cgc_launch(rs_cgc_pos_kernel<DM>, a, (KD + 2 * a->K) * sizeof(float), cgc_grid(a->P), s, "rs_cgc_pos_kernel"))
        return e;
    size_t smem = (KD + 3 * a->K) * sizeof(float);
    if (a->accumulate_grad)
        smem += (size_t)CGC_TILE * (a->D + 1 + a->K + 1) * sizeof(float);
    const int64_t tiles = (a->P + CGC_TILE - 1) / CGC_TILE;
    const int64_t grid = tiles < (int64_t)rs_num_sms() * 4 ? tiles : (int64_t)rs_num_sms() * 4;
    return cgc_launch(rs_cgc_loss_kernel<DM>, a, smem, (int)(grid < 1 ? 1 : grid), s, "rs_cgc_loss_kernel");
}

// forward: zeroes the workspace, then sums -> temperatures -> loss (ws[last] = loss); with accumulate_grad != 0 the loss pass
// also leaves what rs_cgc_bwd needs
extern "C" int rs_cgc_fwd(const rs_cgc_args *a, rs_stream_t stream) {
    if (int e = cgc_check(a, "rs_cgc_fwd"))
        return e;
    cudaStream_t s = (cudaStream_t)stream;
    RS_CUDA(cudaMemsetAsync(a->ws, 0, rs_cgc_workspace_floats(a->K, a->D) * sizeof(float), s));
    return a->D <= 16 ? cgc_fwd_impl<16>(a, s) : cgc_fwd_impl<32>(a, s);
}

// backward: v_features [P, D] = dL/dx * (*grad_loss), from the workspace of an rs_cgc_fwd run with accumulate_grad != 0
extern "C" int rs_cgc_bwd(const rs_cgc_args *a, rs_stream_t stream) {
    if (int e = cgc_check(a, "rs_cgc_bwd"))
        return e;
    RS_CHECK(a->v_features != nullptr, "rs_cgc_bwd: v_features is required");
    const size_t smem = (2 * (size_t)a->K * a->D + 3 * a->K) * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (a->D <= 16)
        return cgc_launch(rs_cgc_grad_kernel<16>, a, smem, cgc_grid(a->P), s, "rs_cgc_grad_kernel");
    return cgc_launch(rs_cgc_grad_kernel<32>, a, smem, cgc_grid(a->P), s, "rs_cgc_grad_kernel");
}
