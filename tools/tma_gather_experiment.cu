// Staging experiment for the compositing producer (VERDICT r1 item 5 / north_star "TMA bulk staging of per-tile splat
// batches"): fetch 32-byte splat records by sorted flatten id into a shared-memory ring, one producer warp per CTA, either
//   A  with two 16-byte cp.async (LDGSTS) per record and lane -- what rs_raster_fwd_kernel does today -- or
//   B  with cp.async.bulk.tensor.2d ... tile::gather4 : ONE instruction fetches the four records of four ids
//      (rows of a 2-D tensor map [E, 32 B]) and completes on the same mbarrier with a transaction count.
// Stand-alone micro-benchmark (not part of librigidsplat.so): per-tile id lists like a c2 frame (8160 tiles, ~640 ids per
// tile drawn from a window of the depth-sorted records, so the L2 behaviour is comparable), records = 1 M x 32 B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/tma_gather_experiment.cu -o tools/_tmp/tma_gather_experiment -lcuda
//   tools/_tmp/tma_gather_experiment [box_rows=1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                                                          \
    do {                                                                                                               \
        cudaError_t e = (x);                                                                                           \
        if (e != cudaSuccess) {                                                                                        \
            printf("{\"error\": \"%s at line %d: %s\"}\n", #x, __LINE__, cudaGetErrorString(e));                       \
            exit(0);                                                                                                   \
        }                                                                                                              \
    } while (0)

#define BATCH 256
#define STAGES 3

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned n) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, unsigned ph) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ void cp16(void *s, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}

// one warp per CTA; tile t owns ids[off[t] .. off[t+1]); MODE 0 = LDGSTS, 1 = gather4
template <int MODE>
__global__ void __launch_bounds__(32) stage_kernel(const __grid_constant__ CUtensorMap tmap, const float4 *__restrict__ records,
                                                   const int *__restrict__ ids, const int *__restrict__ off,
                                                   float *__restrict__ sums) {
    __shared__ __align__(128) float ring[STAGES][BATCH * 8];
    __shared__ __align__(8) uint64_t full[STAGES];
    const int lane = threadIdx.x;
    const int lo = off[blockIdx.x], hi = off[blockIdx.x + 1];
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s)
            mbar_init(&full[s], MODE == 0 ? 32 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int nb = (hi - lo + BATCH - 1) / BATCH;
    float acc = 0.f;
    // issue up to STAGES batches ahead, then consume (sum one float per record) in order
    auto issue = [&](int b) {
        float *base = ring[b % STAGES];
        const int start = lo + b * BATCH, n = min(BATCH, hi - start);
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < BATCH / 32; ++k) {
                const int t = k * 32 + lane;
                if (t < n) {
                    const int g = ids[start + t];
                    cp16(base + t * 8, records + (size_t)g * 2);
                    cp16(base + t * 8 + 4, records + (size_t)g * 2 + 1);
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];" ::"r"(smem_u32(&full[b % STAGES])) : "memory");
        } else {
            const int groups = (n + 3) / 4; // gather4 instructions of this batch; a ragged tail repeats its last id
            if (lane == 0)
                asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared.b64 st, [%0], %1;\n}" ::"r"(smem_u32(&full[b % STAGES])),
                             "r"(groups * 128)
                             : "memory");
            __syncwarp();
#pragma unroll
            for (int k = 0; k < BATCH / 128; ++k) {
                const int grp = k * 32 + lane;
                if (grp < groups) {
                    int r[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        r[q] = ids[start + min(grp * 4 + q, n - 1)];
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
                                 "%3, %4, %5, %6}], [%7];" ::"r"(smem_u32(base + grp * 32)),
                                 "l"(&tmap), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(smem_u32(&full[b % STAGES]))
                                 : "memory");
                }
            }
        }
    };
    for (int b = 0; b < min(nb, STAGES); ++b)
        issue(b);
    for (int b = 0; b < nb; ++b) {
        while (!mbar_try(&full[b % STAGES], (b / STAGES) & 1)) {
        }
        const float *base = ring[b % STAGES];
        const int n = min(BATCH, hi - (lo + b * BATCH));
        for (int t = lane; t < n; t += 32)
            acc += base[t * 8] + base[t * 8 + 7];
        __syncwarp();
        if (b + STAGES < nb)
            issue(b + STAGES);
    }
    if (MODE == 0)
        asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0)
        sums[blockIdx.x] = acc;
}

int main(int argc, char **argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const int E = 1000000, TILES = 8160;
    std::vector<float> rec((size_t)E * 8);
    srand(1);
    for (auto &v : rec)
        v = (float)(rand() % 1000) * 0.001f;
    std::vector<int> off(TILES + 1, 0), ids;
    for (int t = 0; t < TILES; ++t) {
        const int n = 200 + rand() % 880; // ~640 per tile
        const int window = rand() % (E - 60000);
        for (int i = 0; i < n; ++i)
            ids.push_back(window + rand() % 60000);
        off[t + 1] = (int)ids.size();
    }
    float *d_rec, *d_sums[2];
    int *d_ids, *d_off;
    CK(cudaMalloc(&d_rec, rec.size() * 4));
    CK(cudaMalloc(&d_ids, ids.size() * 4));
    CK(cudaMalloc(&d_off, off.size() * 4));
    for (int m = 0; m < 2; ++m)
        CK(cudaMalloc(&d_sums[m], TILES * 4));
    CK(cudaMemcpy(d_rec, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {8, (cuuint64_t)E};
    cuuint64_t gstride[1] = {32};
    cuuint32_t box[2] = {8, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_rec, gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("{\"error\": \"cuTensorMapEncodeTiled failed with %d (box rows %d)\"}\n", (int)r, box_rows);
        return 0;
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 13; ++rep) {
            if (rep == 3)
                CK(cudaEventRecord(e0));
            if (mode == 0)
                stage_kernel<0><<<TILES, 32>>>(tmap, (const float4 *)d_rec, d_ids, d_off, d_sums[0]);
            else
                stage_kernel<1><<<TILES, 32>>>(tmap, (const float4 *)d_rec, d_ids, d_off, d_sums[1]);
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms[mode], e0, e1));
        ms[mode] /= 10.f;
    }
    std::vector<float> s0(TILES), s1(TILES);
    CK(cudaMemcpy(s0.data(), d_sums[0], TILES * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(s1.data(), d_sums[1], TILES * 4, cudaMemcpyDeviceToHost));
    int same = 0;
    for (int t = 0; t < TILES; ++t)
        same += s0[t] == s1[t];
    printf("{\"experiment\": \"stage 32-byte records by id into a shared-memory ring, one producer warp per CTA, %d tiles, %zu ids\", "
           "\"ldgsts_2x16B_per_record_us\": %.1f, \"tma_gather4_us\": %.1f, \"gather4_over_ldgsts\": %.3f, \"tiles_with_equal_checksum\": %d, "
           "\"box_rows\": %d, \"records_per_us_ldgsts\": %.0f, \"records_per_us_gather4\": %.0f}\n",
           TILES, ids.size(), ms[0] * 1e3, ms[1] * 1e3, ms[1] / ms[0], same, box_rows, ids.size() / (ms[0] * 1e3), ids.size() / (ms[1] * 1e3));
    return 0;
}
