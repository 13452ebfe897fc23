#!/bin/bash
# Is the device code of the working tree the same as that of another revision?  Compiles flux_b200/csrc/*.cu of both
# with the library's flags and compares `cuobjdump -sass`, ignoring what depends on the source path (the file
# identifier and the hash in the names of anonymous namespaces).  No GPU needed.
# usage: tools/sass_identical.sh <git-revision>      exit 0 = every translation unit identical
set -e
cd "$(dirname "$0")/.."
REV=${1:?usage: tools/sass_identical.sh <git-revision>}
T=$(mktemp -d)
mkdir -p $T/old/flux_b200/csrc $T/old/include $T/sass
for f in $(git ls-tree --name-only $REV flux_b200/csrc/); do git show $REV:$f > $T/old/$f; done
git show $REV:include/fluxb200.h > $T/old/include/fluxb200.h
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2"
strip() { cuobjdump -sass $1 | grep -v "identifier\|Function :\|\.section\|Fatbin\|=====" | sed 's/_GLOBAL__N__[0-9a-f]*//g'; }
rc=0
for s in api render render_regen render_wave2 samplegen; do
    (nvcc $F -c -o $T/sass/old_$s.o $T/old/flux_b200/csrc/$s.cu 2>/dev/null; nvcc $F -c -o $T/sass/new_$s.o flux_b200/csrc/$s.cu 2>/dev/null) &
done
wait
for s in api render render_regen render_wave2 samplegen; do
    if [ "$(strip $T/sass/old_$s.o | md5sum)" = "$(strip $T/sass/new_$s.o | md5sum)" ]; then echo "$s.cu: SASS identical to $REV ($(strip $T/sass/new_$s.o | wc -l) lines)"
    else echo "$s.cu: SASS DIFFERS from $REV"; rc=1; fi
done
rm -rf $T
exit $rc
