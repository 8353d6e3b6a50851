#!/bin/bash
# Development aid: standalone attention harness (tools/attn_trace.cu) per macro set -> tools/bin/at_NAME
#   tools/build_attn_variants.sh NAME "-D..." [NAME "-D..." ...]
root=$(cd "$(dirname "$0")/.." && pwd); mkdir -p $root/tools/bin
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr $2 -o $root/tools/bin/at_$1 \
    $root/tools/attn_trace.cu $root/tools/attn_stub.cu $root/wav2vec-s_b200/csrc/layout.cu -lcuda 2>&1 | grep -v "warning\|^$\|declared but never\|\^\|Remark\|^ *[a-z_]* *$" &
  shift 2
done
wait; ls $root/tools/bin
