#!/bin/bash
# Build A/B variants of libfluxb200.so with different compile-time knobs into flux_b200/lib/variants/.
# usage: tools/build_variants.sh "name:-DFOO=1 -DBAR=2" "name2:..."
set -e
cd "$(dirname "$0")/.."
mkdir -p flux_b200/lib/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 -shared -ccbin /usr/bin/g++ $flags \
    -o flux_b200/lib/variants/lib_$name.so flux_b200/csrc/api.cu flux_b200/csrc/render.cu flux_b200/csrc/render_regen.cu flux_b200/csrc/samplegen.cu &
done
wait
ls -la flux_b200/lib/variants/
