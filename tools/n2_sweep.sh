# 2-GPU sweep of the gradient-exchange knobs (bucket sizes, NCCL CTA cap = SMs the persistent kernels leave free)
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-workloads > gpurun_out/r3r_n2_$tag.json 2> gpurun_out/r3r_n2_$tag.err; }
run c8_t1 LASR_TAIL_MB=0.3,1,2,4 NCCL_MAX_CTAS=8 LASR_SM_RESERVE=8
run c16 NCCL_MAX_CTAS=16 LASR_SM_RESERVE=16
run c16_t1 LASR_TAIL_MB=0.3,1,2,4 NCCL_MAX_CTAS=16 LASR_SM_RESERVE=16
run c12 NCCL_MAX_CTAS=12 LASR_SM_RESERVE=12
run c8_res4 NCCL_MAX_CTAS=8 LASR_SM_RESERVE=4
