#!/bin/bash
# Round profile on the GPU box: plain bench (must exit 0) -> ncu launch list of the same command -> optionally one
# `ncu --set full` capture of the dominant render kernel at 4096 spp (a full-set capture replays the kernel ~40
# times: at 16384 spp that is 25 GPU-minutes, at 4096 spp 5).  usage: tools/profile_round.sh <tag> [full]
TAG=${1:-rX}; FULL=${2:-}
python bench.py > gpurun_out/bench_$TAG.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/bench_$TAG.log; exit 1; }
tail -1 gpurun_out/bench_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs c1,c3,c4 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
if [ -n "$FULL" ]; then
  ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -o gpurun_out/full_$TAG \
      python bench.py --root 64 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "full capture rc=$?"; ls -la gpurun_out/full_$TAG.ncu-rep
  # the two BVH kernels (config 5 ray batch, config 3 mesh crop)
  python tools/prof_bvh.py > gpurun_out/prof_bvh_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'trace_rays_bvh|render_regen' -o gpurun_out/bvh_$TAG \
      python tools/prof_bvh.py > gpurun_out/ncu_bvh_$TAG.log 2>&1
  echo "bvh capture rc=$?"; ls -la gpurun_out/bvh_$TAG.ncu-rep
fi
