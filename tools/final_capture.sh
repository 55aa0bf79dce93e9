set -x
R=${1:-r2e}
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-workloads > gpurun_out/${R}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-workloads > gpurun_out/${R}_ncu.log 2>&1
python tools/prof_step.py context_aishell_b32_16s_bf16 2 > gpurun_out/${R}_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches_cfg4.csv python tools/prof_step.py context_aishell_b32_16s_bf16 2 > gpurun_out/${R}_ncu_cfg4_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ctc_grad_large|log_softmax_fwd|bilstm_fwd3|bilstm_bwd3" -c 4 -o gpurun_out/${R}_cfg4 -f python tools/prof_step.py context_aishell_b32_16s_bf16 1 > gpurun_out/${R}_prof_ncu.log 2>&1
ls -la gpurun_out | tail -12
