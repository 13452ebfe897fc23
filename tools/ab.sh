#!/bin/bash
# Time every variant in flux_b200/lib/variants with a short bench run (on the GPU box).
ROOT=${1:-32}
for so in flux_b200/lib/variants/lib_*.so; do
  FLUXB200_LIB=$PWD/$so python bench.py --root $ROOT --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs none 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', round(d['value'],1), 'Msamples/s  frac', round(d['roofline']['frac'],4))" || echo "$so FAILED"
done
