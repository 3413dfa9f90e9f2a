// Layout of the radix-sort workspace and tile geometry, shared by sort.cu and the kernels that pre-compute histograms
// for it (isect.cu: the emission kernel builds the tile-key histograms while it writes the keys).
#pragma once
#define SORT_THREADS 256
#ifndef SORT_ITEMS
#define SORT_ITEMS 16
#endif
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_WARPS (SORT_THREADS / 32)
#define RADIX_BITS 8
#define RADIX (1 << RADIX_BITS)
#define SORT_MAX_PASSES 8
// workspace layout (uint32 words): [0, 8*256) digit histograms per pass | [2048, 2048+8) tile tickets per pass |
// [2304, ...) look-back words: pass-major, [pass][tile][256]
#define WS_HIST 0
#define WS_TICKET (SORT_MAX_PASSES * RADIX)
#define WS_LOOKBACK (WS_TICKET + 256)

// Digit layout of a sort over key bits [begin_bit, end_bit): the fewest 8-bit-or-narrower passes, with the bits spread
// EVENLY over them (14 bits -> 7 + 7, not 8 + 6; 32 -> 8 x 4).  Narrower digits in every pass mean fewer, longer runs per
// 4096-pair tile when the re-ordered tile is written out, i.e. fuller sectors for the scattered stores.
__host__ __device__ inline int sort_num_passes(int total_bits) { return (total_bits + RADIX_BITS - 1) / RADIX_BITS; }
__host__ __device__ inline int sort_digit_width(int total_bits) {
    const int passes = sort_num_passes(total_bits);
    return passes > 0 ? (total_bits + passes - 1) / passes : RADIX_BITS;
}
