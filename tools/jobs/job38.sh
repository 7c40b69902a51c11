for n in 2 4 8 16; do python tools/timeline_forward.py 256 $n > gpurun_out/tl_n$n.txt 2>&1; head -1 gpurun_out/tl_n$n.txt; done
python -m terrain_diffusion_b200.build > /dev/null
