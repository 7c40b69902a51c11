set -u
python -m terrain_diffusion_b200.build > /dev/null
b() { python bench.py --steps 20 --warmup 5 $2 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; tail -1 gpurun_out/bench_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['frac'], d['e2e']['value'])"; }
b f1 ""
b f1_16 "--tiles 16 --no-cpu-baseline"
b f1_512 "--size 512 --no-cpu-baseline"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/profile_round.sh r02f > gpurun_out/profile_round_r02f.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
