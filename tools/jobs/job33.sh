b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tiles ${2:-1} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'tiles ${2:-1}', round(d['value'],1), round(d['ms_per_step'],4))"; }
timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -2
b sep 1; b sep 16
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:igemm_kernel -s 153 -c 2 --csv python bench.py --steps 2 --warmup 1 --tiles 16 --no-cpu-baseline 2>/dev/null | grep -i "igemm" | cut -d, -f5,12- | head -8
cp /dev/null /dev/null
