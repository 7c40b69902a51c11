timeout 900 python -m pytest tests/test_lazy_canvas_gpu.py tests/test_stages_gpu.py tests/test_multiphase_gpu.py -x -q 2>&1 | tail -3
python bench.py --workload world --steps 10 --warmup 3 > gpurun_out/bench_world.json 2> gpurun_out/bench_world.err; tail -1 gpurun_out/bench_world.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('world', d['ttft'], d['ttst'], d['peak_vram_mb'])"; tail -3 gpurun_out/bench_world.err
timeout 600 python tools/profile_world_get.py > gpurun_out/world_profile2.txt 2>&1; head -30 gpurun_out/world_profile2.txt | cut -c1-150
