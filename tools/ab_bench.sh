#!/bin/bash
# Same-box A/B: alternate the product library and a variant, print ms/step and the GEMM / attention device times.
#   tools/ab_bench.sh VARIANT_NAME [ROUNDS]
v=$1; n=${2:-2}
for i in $(seq $n); do
  for lib in "" "wav2vec-s_b200/lib/libw2vs_$v.so"; do
    W2VS_LIBRARY=$lib timeout 200 python bench.py --steps 8 --kernel-detail --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_detail']
g={n.split('[')[1][:-1]:v[0] for n,v in k.items() if 'M=95744' in n}
print('${lib:-product}'.split('/')[-1].ljust(22), 'ms/step %.2f' % d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], d['kernel_ms_per_step'], ' '.join('%s=%.2f' % (a.split(',',3)[1]+a.split(',',3)[2]+('g' if 'gelu' in a else ''), b) for a,b in g.items()))"
  done
done
