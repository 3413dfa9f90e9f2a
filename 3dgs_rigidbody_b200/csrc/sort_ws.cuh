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
