#!/bin/bash
# Build A/B variants of libfluxb200.so with different compile-time knobs into flux_b200/lib/variants/.
# usage: tools/build_variants.sh "name:-DFOO=1 -DBAR=2" "name2:..."   (then tools/ab.sh on the GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p flux_b200/lib/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  python flux_b200/build.py -o flux_b200/lib/variants/lib_$name.so $flags > /dev/null &
done
wait
ls -la flux_b200/lib/variants/
