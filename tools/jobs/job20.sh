b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tiles ${2:-1} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'tiles ${2:-1}', round(d['value'],1), round(d['ms_per_step'],4))"; }
for w in 0 1 2 0; do export TDX_NVCC_DEFINES="TDX_V_WAIT=$w"; python -m terrain_diffusion_b200.build >/dev/null; b "wait=$w" 1; b "wait=$w" 16; done
export TDX_NVCC_DEFINES="TDX_V_WAIT=1 TDX_V_GUARD=0"; python -m terrain_diffusion_b200.build >/dev/null; b "wait=1,guard=0" 1; b "wait=1,guard=0" 16
export TDX_NVCC_DEFINES="TDX_V_WAIT=1"; python -m terrain_diffusion_b200.build >/dev/null; TDX_CVEC_HALF=0 b "wait=1,half=0" 1;  TDX_CVEC_HALF=0 b "wait=1,half=0" 16
