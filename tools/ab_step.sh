#!/bin/bash
# A/B of one environment switch on the headline step (device-resident ms/step, 3 runs each, interleaved)
# usage: tools/ab_step.sh VAR val_a val_b
var=$1; a=$2; b=$3
for i in 1 2 3; do
  for v in $a $b; do
    env $var=$v python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('$var=$v', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), {k:f[k]['ms'] for k in ('pwconv_gemm','bn_pass','dwconv','ctc_fwd')})
"
  done
done
