#!/bin/bash
# Same-box comparison of several variant libraries: per-class device ms of one cfg3 step.
#   tools/ab_libs.sh NAME1 NAME2 ...   ("product" = wav2vec-s_b200/lib/libw2vs.so)
for v in "$@"; do
  lib=""; [ "$v" != product ] && lib="wav2vec-s_b200/lib/libw2vs_$v.so"
  W2VS_LIBRARY=$lib timeout 200 python bench.py --steps 6 --no-cpu-baseline --no-incremental 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v'.ljust(12), 'ms/step %.2f' % d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], d['kernel_ms_per_step'])"
done
