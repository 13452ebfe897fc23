#!/bin/bash
# A/B of BVH traversal variants on the GPU box: the shipped library, then every flux_b200/lib/variants/lib_*.so,
# on config 5 (20 M rays, best of 4 launches per chunk) and config 3 (full frame).  usage: tools/ab_bvh.sh [tag]
TAG=${1:-ab}
mkdir -p gpurun_out
LOG=gpurun_out/ab_bvh_$TAG.log
echo base > $LOG
timeout 300 python tools/bench_configs.py ${AB_CONFIGS:-c5q c3} >> $LOG 2>&1
for so in flux_b200/lib/variants/lib_*.so; do echo $so >> $LOG; FLUXB200_LIB=$PWD/$so timeout 300 python tools/bench_configs.py ${AB_CONFIGS:-c5q c3} >> $LOG 2>&1; done
grep -o '^base\|^flux.*so\|"Mrays_per_s": [0-9.]*\|"Msamples_per_s": [0-9.]*\|bvh_bitwise_equal": [a-z]*' $LOG
