// Asynchronous-copy and mbarrier helpers shared by the compositing kernels (raster_fwd.cu, raster_bwd.cu).
#pragma once
#include "common.cuh"

__device__ __forceinline__ void rs_cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void rs_cp_async4(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
// waits for ALL of this thread's cp.async copies, committed to a group or not (the ring tracks its copies through
// cp.async.mbarrier.arrive and never commits groups, so wait_group would not cover them)
__device__ __forceinline__ void rs_cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void rs_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void rs_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- mbarrier helpers (shared::cta) --------------------------------------------------------------------------------
__device__ __forceinline__ unsigned rs_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(rs_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool rs_mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(rs_smem_addr(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void rs_mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared.b64 st, [%0];\n\t}\n" ::"r"(rs_smem_addr(bar)) : "memory");
}
// arrives on `bar` once all cp.async copies issued so far by this thread have landed (does not bump the pending count)
__device__ __forceinline__ void rs_cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(rs_smem_addr(bar)) : "memory");
}

// explicit shared-space loads from 32-bit shared addresses (keeps the generic->shared conversion out of the inner loop)
__device__ __forceinline__ float4 rs_lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int rs_lds32i(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}

#ifdef RS_RASTER_STATS
__device__ unsigned long long rs_stats[8];
extern "C" void rs_raster_stats(unsigned long long *out) { // {iterations, with >= 1 passing lane, passing, active, chunks}
    cudaMemcpyFromSymbol(out, rs_stats, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(rs_stats, z, sizeof(z));
}
#endif

// packs the staging records [n_rows, 8] from means2d / conics / opacities (defined in raster_fwd.cu)
int rs_raster_pack_records(const rs_raster_fwd_args &a, cudaStream_t s);
