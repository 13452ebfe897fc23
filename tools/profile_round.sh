#!/bin/bash
# Round profile on the GPU box: plain bench (must exit 0) -> ncu launch list of the same command -> one
# `ncu --set full` capture of the dominant render kernel.  usage: tools/profile_round.sh <tag> [bench args...]
TAG=${1:-rX}; shift
ARGS=${@:---steps 2 --warmup 3}
python bench.py $ARGS > gpurun_out/bench_$TAG.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/bench_$TAG.log; exit 1; }
tail -1 gpurun_out/bench_$TAG.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -o gpurun_out/full_$TAG \
    python bench.py $ARGS --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/full_$TAG.ncu-rep
