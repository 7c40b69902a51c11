set -u
b() { python bench.py --steps 20 --warmup 5 $2 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; tail -1 gpurun_out/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['frac'], d['e2e']['value'])"; }
b b1 ""
b b1_16 "--tiles 16 --no-cpu-baseline"
b b1_512 "--size 512 --no-cpu-baseline"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm_kernel -s 153 -c 2 -f -o gpurun_out/r02b_enc0 python bench.py --steps 2 --warmup 1 --tiles 16 --no-cpu-baseline > gpurun_out/ncu_enc0.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm_kernel -s 218 -c 2 -f -o gpurun_out/r02b_dec0 python bench.py --steps 2 --warmup 1 --tiles 16 --no-cpu-baseline > gpurun_out/ncu_dec0.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
