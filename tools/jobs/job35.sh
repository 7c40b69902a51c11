b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tiles ${2:-1} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'tiles ${2:-1}', round(d['value'],1), round(d['ms_per_step'],4))"; }
for v in 0 1 0 1; do export TDX_NVCC_DEFINES="TDX_V_RES0_KERNEL=$v"; python -m terrain_diffusion_b200.build >/dev/null; b "res0k=$v" 1; b "res0k=$v" 16; done
export TDX_NVCC_DEFINES=""; python -m terrain_diffusion_b200.build >/dev/null
timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_unet_gpu.py tests/test_parity_r2_gpu.py -x -q 2>&1 | tail -2
