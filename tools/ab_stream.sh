#!/bin/bash
for i in 1 2; do
for lib in "" "wav2vec-s_b200/lib/libw2vs_ns2.so"; do
  for wl in stream_large_b16 stream_large_b1; do
    W2VS_LIBRARY=$lib timeout 120 python bench.py --workload $wl --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${lib:-product}'.split('/')[-1].ljust(20), '$wl', 'p50 %.4f p90 %.4f' % (d['value'], d['p90_ms']), d['clocks']['sm_mhz'])"
  done
done
done
