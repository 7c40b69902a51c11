b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tiles ${2:-1} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'tiles ${2:-1}', round(d['value'],1), round(d['ms_per_step'],4))"; }
python -m terrain_diffusion_b200.build >/dev/null; b "ahead" 1; b "ahead" 16
timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_unet_gpu.py tests/test_parity_r2_gpu.py -x -q 2>&1 | tail -2
python tools/timeline_forward.py 256 16 > gpurun_out/tl16_ahead.txt 2>&1
TDX_NVCC_DEFINES="TDX_EPI_WQ=4 TDX_EPI_CHUNK=16" python tools/timeline_forward.py 256 16 > gpurun_out/tl16_wq4.txt 2>&1
tail -c 300 gpurun_out/tl16_wq4.txt
