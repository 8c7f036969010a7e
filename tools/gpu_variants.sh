#!/bin/bash
# usage (GPU box): tools/gpu_variants.sh lib1.so lib2.so ...   -- bench each differently-tuned build of the same sources
for so in "$@"; do
  echo "== $so"
  SS_B200_LIB=$so python bench.py --steps 128 --warmup 16 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3fM ms/step %.4f e2e %.3fM aba_ms %.4f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline_aba']['avg_launch_ms']))"
done
