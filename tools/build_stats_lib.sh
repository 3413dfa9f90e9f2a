#!/bin/bash
# -DRS_RASTER_STATS build of librigidsplat (instrumentation only) -> tools/_tmp/librigidsplat_stats.so
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/statlib tools/_tmp
for f in abi project project_bwd isect sort raster_fwd raster_bwd frame sh exchange cgc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -use_fast_math -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DRS_RASTER_STATS -c 3dgs_rigidbody_b200/csrc/$f.cu -o /tmp/statlib/$f.o &
done
wait
nvcc -shared -o tools/_tmp/librigidsplat_stats.so /tmp/statlib/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
