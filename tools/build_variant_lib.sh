#!/bin/bash
# Builds a variant of librigidsplat.so with extra nvcc flags for A/B experiments (never the shipped library):
#   tools/build_variant_lib.sh items8 -DSORT_ITEMS=8
#   RIGIDSPLAT_LIB=tools/_tmp/librigidsplat_items8.so python bench.py --no-extras
# Known switches: -DSORT_ITEMS=8 (2048-pair sort tiles), -DSORT_LB_MODE=1|2 (look-back timing experiments, MODE 1 sorts
# wrongly on purpose), -DRS_RASTER_STATS (cull-test counters, see tools/raster_stats.py).
set -e
cd "$(dirname "$0")/.."
name="$1"; shift
[ -n "$name" ] || { echo "usage: $0 NAME [nvcc flags...]"; exit 2; }
obj="/tmp/rs_variant_$name"; mkdir -p "$obj" tools/_tmp
for f in 3dgs_rigidbody_b200/csrc/*.cu; do
  b=$(basename "$f" .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -use_fast_math -lineinfo -std=c++17 -Xcompiler -fPIC \
       --expt-relaxed-constexpr "$@" -c "$f" -o "$obj/$b.o" &
done
wait
nvcc -shared -o "tools/_tmp/librigidsplat_$name.so" "$obj"/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo "tools/_tmp/librigidsplat_$name.so"
