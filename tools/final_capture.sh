set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_reference_arm.json 2> gpurun_out/r2d_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-workloads > gpurun_out/r2d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-workloads > gpurun_out/r2d_ncu.log 2>&1
python tools/prof_step.py asr13x1_b32_16s_bf16 2 > gpurun_out/r2d_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ctc_lattice_warp2|bn_bwd_apply_ring|ctc_grad" -c 6 -o gpurun_out/r2d_prof_ctc_bnring -f python tools/prof_step.py asr13x1_b32_16s_bf16 2 > gpurun_out/r2d_prof_ncu.log 2>&1
ls -la gpurun_out | tail -8
