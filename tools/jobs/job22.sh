b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tiles ${2:-1} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'tiles ${2:-1}', round(d['value'],1), round(d['ms_per_step'],4))"; }
export TDX_NVCC_DEFINES="TDX_EPI_WQ=4 TDX_EPI_CHUNK=16"; python -m terrain_diffusion_b200.build >/dev/null; b "wq4ch16" 1; b "wq4ch16" 16
timeout 600 python -m pytest tests/test_igemm_gpu.py -x -q 2>&1 | tail -2
export TDX_NVCC_DEFINES="TDX_EPI_WQ=4 TDX_EPI_CHUNK=32"; python -m terrain_diffusion_b200.build >/dev/null; b "wq4ch32" 1; b "wq4ch32" 16
export TDX_NVCC_DEFINES=""; python -m terrain_diffusion_b200.build >/dev/null; b "base" 1; b "base" 16
