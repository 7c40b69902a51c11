timeout 1500 python tools/tune_igemm.py decoder > gpurun_out/tune_graph.log 2>&1; tail -12 gpurun_out/tune_graph.log
cp terrain_diffusion_b200/tuned_shapes.json gpurun_out/tuned_shapes_graph.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a7.json 2>gpurun_out/bench_a7.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a7.json"))
print("1 tile", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python bench.py --steps 20 --warmup 5 --tiles 16 --no-cpu-baseline > gpurun_out/bench_a7_16.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a7_16.json"))
print("16 tiles", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python bench.py --steps 20 --warmup 5 --tiles 1 --size 512 --no-cpu-baseline > gpurun_out/bench_a7_512.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a7_512.json"))
print("1 x 512", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
