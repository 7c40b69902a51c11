b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$2', round(d['value'],1), round(d['ms_per_step'],4))"; }
for mb in 0 32 48 64 96 0; do export TDX_L2_PERSIST_MB=$mb; b "persist=$mb" "--tiles 1"; b "persist=$mb" "--tiles 16"; b "persist=$mb" "--size 512"; done
