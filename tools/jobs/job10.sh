timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a5.json 2>gpurun_out/bench_a5.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a5.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python tools/timeline_forward.py 256 16 > gpurun_out/timeline_a5_16.txt 2>&1; head -1 gpurun_out/timeline_a5_16.txt; python bench.py --impl reference-gpu --steps 20 --warmup 3 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; cat gpurun_out/bench_refgpu.json | cut -c1-1500; tail -3 gpurun_out/bench_refgpu.err; python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels.txt 2>&1; cat gpurun_out/hbm_kernels.txt; timeout 900 python tools/tune_igemm.py decoder > gpurun_out/tune.log 2>&1; tail -3 gpurun_out/tune.log; cp terrain_diffusion_b200/tuned_shapes.json gpurun_out/tuned_shapes.json; python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a5t.json 2>gpurun_out/bench_a5t.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a5t.json"))
print("tuned", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python tools/timeline_forward.py 256 16 > gpurun_out/timeline_a5t_16.txt 2>&1; head -1 gpurun_out/timeline_a5t_16.txt
