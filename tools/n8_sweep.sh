#!/bin/bash
# 8-GPU sweep of the gradient-exchange knobs (NCCL algorithm, CTA cap = SMs left to NCCL, bucket grading)
# usage (8-GPU box): bash tools/n8_sweep.sh <prefix>
pre=${1:-r2d}
run() { tag=$1; shift; timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-workloads > gpurun_out/${pre}_n8_$tag.json 2> gpurun_out/${pre}_n8_$tag.err; tail -c 300 gpurun_out/${pre}_n8_$tag.json | head -c 10 > /dev/null; }
run base NCCL_DEBUG=WARN
run nvls NCCL_ALGO=NVLS
run nvls4 NCCL_ALGO=NVLS NCCL_MAX_CTAS=4 LASR_SM_RESERVE=4
run fewbuckets LASR_TAIL_MB=1.5,6
for f in gpurun_out/${pre}_n8_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), round(d['value']), d.get('allreduce_exposed_us'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
P
done
