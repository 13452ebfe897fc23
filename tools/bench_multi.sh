#!/bin/bash
# bench.py on N GPUs of one box, as the driver launches it (torchrun, one rank per GPU); with "ab" also the A/B of the
# frame assembly (peer memory vs NCCL all_gather) and of 4-row tiles.  usage: tools/bench_multi.sh N tag [ab]
N=${1:-2}; TAG=${2:-multi}; AB=${3:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_g$N.json 2> gpurun_out/bench_${TAG}_g$N.err
if [ -n "$AB" ]; then
  $TR bench.py --gpus $N --steps 3 --warmup 3 --configs none --gather nccl > gpurun_out/bench_${TAG}_g${N}_nccl.json 2>> gpurun_out/bench_${TAG}_g$N.err
  $TR bench.py --gpus $N --steps 3 --warmup 3 --configs none --gather peer --tile-rows 4 > gpurun_out/bench_${TAG}_g${N}_tile4.json 2>> gpurun_out/bench_${TAG}_g$N.err
fi
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_${TAG}_g$N.err | tail -c 1500
for f in gpurun_out/bench_${TAG}_g$N*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "assembly_ms", round(d["assembly_ms_per_step"], 3),
          "kernel max/min", round(d["kernel_ms_max_over_ranks"], 2), round(d["kernel_ms_min_over_ranks"], 2), d["frame_sha256"][:12], d["frame_assembly"])
    for k, v in (d.get("configs") or {}).items():
        if isinstance(v, dict):
            print("  ", k, round(v["value"], 1), v["unit"], "resident", round(v["value_resident"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
