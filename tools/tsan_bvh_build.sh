#!/bin/bash
# The BVH builder (host threads: slices, nested partitions, parallel node emission; arrays that are not zeroed before
# they are filled) under ThreadSanitizer — or, with `asan`, AddressSanitizer + UBSan — no GPU: 300 K random triangles +
# 3 spheres (one oversized) built twice; prints the sizes and whether the two trees are equal.  Any "WARNING:
# ThreadSanitizer" / "ERROR: AddressSanitizer" / "runtime error" line is a failure.
# usage: tools/tsan_bvh_build.sh [asan]   (round 2: both clean, same 1)
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
if [ "$1" = asan ]; then S="-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined"; L="-fsanitize=address,undefined"
else S="-Xcompiler -fsanitize=thread"; L="-fsanitize=thread"; fi
F="-std=c++17 -O1 -g -Iflux_b200/csrc $S -Xcompiler -fno-omit-frame-pointer"
nvcc $F -c tools/tsan_bvh_main.cu -o $T/main.o
nvcc $F -c flux_b200/csrc/bvh_build.cu -o $T/bvh.o
g++ $L -o $T/san_bvh $T/main.o $T/bvh.o -L/usr/local/cuda/lib64 -lcudart -lpthread
TSAN_OPTIONS="halt_on_error=0" ASAN_OPTIONS="detect_leaks=0" $T/san_bvh
rm -rf $T
