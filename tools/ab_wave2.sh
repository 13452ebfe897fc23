#!/bin/bash
# A/B of render_wave2 variants on the GPU box: the shipped library, then every flux_b200/lib/variants/lib_*.so, on the
# headline scene at sample_root ROOT (default 64 = 4096 spp).  usage: tools/ab_wave2.sh [tag] [root]
TAG=${1:-ab}; ROOT=${2:-64}
mkdir -p gpurun_out
LOG=gpurun_out/ab_wave2_$TAG.log
: > $LOG
run() {
  python bench.py --root $ROOT --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs ${AB_CONFIGS:-none} 2>&1 | grep '^{' | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d.get('configs') or {}; print('$1', round(d['value'],1), 'Msamples/s  frac', round(d['roofline']['frac'],4), d['frame_sha256'][:10], {k:round(v['value_resident'],1) for k,v in c.items() if isinstance(v,dict)})" || echo "$1 FAILED"
}
run base | tee -a $LOG
for so in flux_b200/lib/variants/lib_*.so; do FLUXB200_LIB=$PWD/$so run $so | tee -a $LOG; done
