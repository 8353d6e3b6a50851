#!/bin/bash
# Development aid: build a second libw2vs (kernel variant) next to the product library for same-box A/B timing.
#   tools/build_variant.sh NAME "-DMACRO=1 ..."   ->  wav2vec-s_b200/lib/libw2vs_NAME.so ; run with W2VS_LIBRARY=<that path>
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/wav2vec-s_b200/build/variant_$name
mkdir -p $out
for f in $root/wav2vec-s_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $defs -c $f -o $out/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -o $root/wav2vec-s_b200/lib/libw2vs_$name.so $out/*.o -gencode arch=compute_100a,code=sm_100a
echo $root/wav2vec-s_b200/lib/libw2vs_$name.so
