#!/bin/bash
# Light ncu capture of one render kernel launch (instruction counts per SASS line + scheduler/warp-state
# stats), cheap enough to run after every kernel change.  usage (on the GPU box): tools/prof_light.sh <tag> [root]
TAG=${1:-x}; ROOT=${2:-16}
CMD="python bench.py --root $ROOT --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
python - <<PY
import json
d=json.loads(open("gpurun_out/plain_$TAG.log").read().strip().splitlines()[-1])
print("plain:", round(d["value"],1), "Msamples/s  frac", round(d["roofline"]["frac"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],2))
PY
ncu --section SourceCounters --section InstructionStats --section WarpStateStats --section SchedulerStats \
    --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section SpeedOfLight \
    --clock-control none --import-source on -k regex:render_ -s ${SKIP:-3} -c 1 -o gpurun_out/light_$TAG $CMD > gpurun_out/ncu_light_$TAG.log 2>&1
tail -2 gpurun_out/ncu_light_$TAG.log
