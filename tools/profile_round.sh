#!/bin/bash
# Round profile on the GPU box (run via gpurun): launch list of the bench command, DRAM bytes of every igemm launch of
# one forward, and a full-section capture of three representative igemm launches.  Outputs land in gpurun_out/.
set -u
K='regex:igemm_kernel|conv_in_kernel|im2col_in_kernel|conv_out_kernel|embed_kernel|attn_kernel|sched_step_kernel'
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:igemm_kernel -s 75 -c 75 --csv --log-file gpurun_out/igemm_dram.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm_kernel -s 75 -c 6 -f \
  -o gpurun_out/igemm_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
