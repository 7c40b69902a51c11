python bench.py --workload world --steps 10 --warmup 3 > gpurun_out/bench_world.json 2> gpurun_out/bench_world.err; tail -c 1800 gpurun_out/bench_world.json; tail -5 gpurun_out/bench_world.err
