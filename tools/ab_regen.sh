#!/bin/bash
# A/B of the regeneration kernel's linear scan (64 <= spp < 256 on sphere/plane scenes): demo2 at sample_root 12 with kernel mode 2.
TAG=${1:-ab}; LOG=gpurun_out/ab_regen_$TAG.log; mkdir -p gpurun_out; : > $LOG
for so in "" flux_b200/lib/variants/lib_*.so; do
  FLUXB200_LIB=${so:+$PWD/$so} python bench.py --root ${ROOT:-12} --kernel-mode 2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none 2>&1 | grep '^{' | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${so:-base}', round(d['value'],1), 'Msamples/s  frac', round(d['roofline']['frac'],4))" | tee -a $LOG
done
