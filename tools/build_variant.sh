#!/bin/bash
# Build libstorm_b200.so of another git revision (or of the working tree: rev = WORK) next to the current one,
# for same-box A/B timing (tools/ab_libs.py):   tools/build_variant.sh <rev|WORK> <name>
set -eu
rev=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
if [ "$rev" = WORK ]; then
  mkdir -p $tmp/stormbitmaps_b200 && cp -r $root/stormbitmaps_b200/csrc $tmp/stormbitmaps_b200/ && cp -r $root/include $tmp/
else
  git -C $root archive $rev stormbitmaps_b200/csrc include | tar -x -C $tmp
fi
objs=""
for f in $tmp/stormbitmaps_b200/csrc/*.cu; do
  o=$tmp/$(basename $f .cu).o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${NVCC_EXTRA:-} -I $tmp/include -I $tmp/stormbitmaps_b200/csrc -c $f -o $o &
  objs="$objs $o"
done
wait
mkdir -p $root/stormbitmaps_b200/_variants
nvcc -shared -o $root/stormbitmaps_b200/_variants/lib_$name.so $objs -gencode arch=compute_100a,code=sm_100a
rm -rf $tmp
echo $root/stormbitmaps_b200/_variants/lib_$name.so
