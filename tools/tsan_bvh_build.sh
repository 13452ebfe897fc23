#!/bin/bash
# The BVH builder (host threads: slices, nested partitions, parallel node emission) under ThreadSanitizer, no GPU:
# 300 K random triangles + 3 spheres (one oversized) built twice; prints the sizes and whether the two trees are equal.
# Any "WARNING: ThreadSanitizer" line is a failure.  usage: tools/tsan_bvh_build.sh   (round 2: clean, same 1)
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
F="-std=c++17 -O1 -g -Iflux_b200/csrc -Xcompiler -fsanitize=thread -Xcompiler -fno-omit-frame-pointer"
nvcc $F -c tools/tsan_bvh_main.cu -o $T/main.o
nvcc $F -c flux_b200/csrc/bvh_build.cu -o $T/bvh.o
g++ -fsanitize=thread -o $T/tsan_bvh $T/main.o $T/bvh.o -L/usr/local/cuda/lib64 -lcudart -lpthread
TSAN_OPTIONS="halt_on_error=0" $T/tsan_bvh
rm -rf $T
