b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],4))"; }
b "all-on(a9)"
TDX_CVEC_HALF=0 b "half-off"
export TDX_NVCC_DEFINES="TDX_V_GUARD=0"; python -m terrain_diffusion_b200.build >/dev/null; b "guard-off"
export TDX_NVCC_DEFINES="TDX_V_SLIMWAIT=0"; python -m terrain_diffusion_b200.build >/dev/null; b "slimwait-off"
export TDX_NVCC_DEFINES="TDX_V_SLIMWAIT=0 TDX_V_GUARD=0"; python -m terrain_diffusion_b200.build >/dev/null; TDX_CVEC_HALF=0 b "all-off(=a8?)"
export TDX_NVCC_DEFINES="TDX_V_SLIMWAIT=0 TDX_V_GUARD=0"; TDX_CVEC_HALF=1 b "only-half"
unset TDX_NVCC_DEFINES; python -m terrain_diffusion_b200.build >/dev/null; b "all-on again"
