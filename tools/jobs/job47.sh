timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm_kernel -s 153 -c 2 -f -o gpurun_out/r02c_enc0_1tile python bench.py --steps 2 --warmup 1 --tiles 1 --no-cpu-baseline > gpurun_out/ncu_enc0_1tile.log 2>&1
ls -la gpurun_out/r02c_enc0_1tile.ncu-rep
