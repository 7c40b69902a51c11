# the round's closing measurement on one box: GPU tests, every bench line kept under profiles/, the ncu summaries
set -u
python -m terrain_diffusion_b200.build > /dev/null
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
b() { python bench.py $2 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; grep '^{' gpurun_out/bench_$1.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],2), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))"; }
b tiles1 "--steps 20 --warmup 5"
b tiles16 "--steps 20 --warmup 5 --tiles 16 --no-cpu-baseline"
b 1x512 "--steps 20 --warmup 5 --size 512 --no-cpu-baseline"
b latent_stage_b16 "--workload latent --steps 20 --warmup 5"
b world_latency "--workload world --steps 20 --warmup 0"
bash tools/profile_round.sh r02f > gpurun_out/profile_round_r02f.log 2>&1
ls gpurun_out | grep r02f
