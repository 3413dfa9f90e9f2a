#!/bin/bash
# Builds the test-side cub::DeviceRadixSort comparison arm of tools/sort_bench.py (never linked into the product).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_tmp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared tools/cub_sort_ref.cu -o tools/_tmp/libcubsort.so
echo tools/_tmp/libcubsort.so
