b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tiles ${2:-1} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'tiles ${2:-1}', round(d['value'],1), round(d['ms_per_step'],4))"; }
for v in 1 0 1 0; do export TDX_NVCC_DEFINES="TDX_V_TWO_KERNELS=$v"; python -m terrain_diffusion_b200.build >/dev/null; b "two=$v" 1; b "two=$v" 16; done
export TDX_NVCC_DEFINES=""; python -m terrain_diffusion_b200.build >/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
