timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_lazy_canvas_gpu.py tests/test_blend_gpu.py -x -q 2>&1 | tail -3
python bench.py --workload world --steps 10 --warmup 3 > gpurun_out/bench_world.json 2> gpurun_out/bench_world.err; tail -1 gpurun_out/bench_world.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('world', d['ttft'], d['ttst'])"
python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiles1', d['value'], d['e2e']['value'])"
