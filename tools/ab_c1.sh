#!/bin/bash
# A/B of the direct kernel (render.cu) on config 1 (demo1 512x512 @16 spp) and on demo2 at 16 spp: shipped library, then variants.
TAG=${1:-ab}; LOG=gpurun_out/ab_c1_$TAG.log; mkdir -p gpurun_out; : > $LOG
for so in "" flux_b200/lib/variants/lib_*.so; do
  echo "lib: ${so:-base}" | tee -a $LOG
  FLUXB200_LIB=${so:+$PWD/$so} python tools/bench_configs.py c1 2>&1 | grep -o '"Msamples_per_s": [0-9.]*' | tee -a $LOG
done
