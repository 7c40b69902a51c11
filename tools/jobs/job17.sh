timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a9.json 2>gpurun_out/bench_a9.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a9.json"))
print("1 tile", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
PY
python bench.py --steps 20 --warmup 5 --tiles 16 --no-cpu-baseline > gpurun_out/bench_a9_16.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a9_16.json"))
print("16 tiles", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python bench.py --steps 20 --warmup 5 --tiles 1 --size 512 --no-cpu-baseline > gpurun_out/bench_a9_512.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a9_512.json"))
print("1 x 512", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
