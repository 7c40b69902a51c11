python bench.py --workload latent --steps 20 --warmup 5 > gpurun_out/bench_latent16.json 2> gpurun_out/bench_latent16.err; tail -c 2500 gpurun_out/bench_latent16.json; tail -5 gpurun_out/bench_latent16.err
python bench.py --workload latent --tiles 4 --steps 20 --warmup 5 > gpurun_out/bench_latent4.json 2> gpurun_out/bench_latent4.err; tail -c 600 gpurun_out/bench_latent4.json
