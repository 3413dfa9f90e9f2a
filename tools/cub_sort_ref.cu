// Test-side comparison arm for tools/sort_bench.py: cub::DeviceRadixSort::SortPairs, the call the reference makes at
// gsplat/cuda/csrc/IntersectTile.cu:296-339 (64-bit isect ids, int32 flatten ids, DoubleBuffer, bits [0, end_bit)).
// NOT part of librigidsplat.so and never on the product path -- it is the bar the hand-written sort is measured against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared tools/cub_sort_ref.cu -o tools/_tmp/libcubsort.so
#include <cub/cub.cuh>
#include <stdint.h>

template <typename KeyT>
static int sort_impl(void *temp, size_t *temp_bytes, KeyT *ka, KeyT *kb, int32_t *va, int32_t *vb, int64_t n, int begin_bit,
                     int end_bit, void *stream, int *selector) {
    cub::DoubleBuffer<KeyT> dk(ka, kb);
    cub::DoubleBuffer<int32_t> dv(va, vb);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, dk, dv, n, begin_bit, end_bit, (cudaStream_t)stream);
    if (selector)
        *selector = dk.selector;
    return (int)e;
}

extern "C" int cub_sort_pairs_u32(void *temp, size_t *temp_bytes, uint32_t *ka, uint32_t *kb, int32_t *va, int32_t *vb,
                                  int64_t n, int begin_bit, int end_bit, void *stream, int *selector) {
    return sort_impl<uint32_t>(temp, temp_bytes, ka, kb, va, vb, n, begin_bit, end_bit, stream, selector);
}
extern "C" int cub_sort_pairs_u64(void *temp, size_t *temp_bytes, uint64_t *ka, uint64_t *kb, int32_t *va, int32_t *vb,
                                  int64_t n, int begin_bit, int end_bit, void *stream, int *selector) {
    return sort_impl<uint64_t>(temp, temp_bytes, ka, kb, va, vb, n, begin_bit, end_bit, stream, selector);
}
