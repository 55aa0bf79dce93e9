run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus 2 --steps 20 --warmup 5 --no-workloads > gpurun_out/r2l_n2_$tag.json 2> gpurun_out/r2l_n2_$tag.err; }
run tail LASR_X=1
run uniform LASR_TAIL_MB=
run tail_cta8 NCCL_MAX_CTAS=8
run tail_cta4 NCCL_MAX_CTAS=4
run tail_cta2 NCCL_MAX_CTAS=2
python -m pytest tests/test_ddp_gpu.py -m gpu -q --tb=short 2>&1 | tail -3 > gpurun_out/r2l_ddp.log
