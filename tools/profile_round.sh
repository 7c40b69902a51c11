#!/bin/bash
# Round profile on the GPU box (run via gpurun): launch list of the bench command, DRAM + L2 bytes of every igemm launch
# of one forward, and a full-section capture of representative igemm launches (tensor-pipe utilisation included).
# Outputs land in gpurun_out/ (copy the summaries into profiles/ with tools/summarize_profile.py).
set -u
R=${1:-r02}
K='regex:igemm_kernel|conv_in_kernel|im2col_in_kernel|conv_out_kernel|embed_kernel|attn_kernel|sched_step_kernel'
mkdir -p gpurun_out
# 20-step solve = 1 embed + 20 x (im2col + 76 igemm + conv_out): skip the warm-up solves, list ~5 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 400 -c 400 --csv \
  --log-file gpurun_out/${R}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:igemm_kernel -s 400 -c 76 --csv --log-file gpurun_out/${R}_igemm_per_launch.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:igemm_kernel -s 401 -c 6 -f \
  -o gpurun_out/${R}_igemm_full python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -6
