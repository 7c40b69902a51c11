timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a6.json 2>gpurun_out/bench_a6.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a6.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python tools/timeline_forward.py 256 16 > gpurun_out/timeline_a6_16.txt 2>&1; head -1 gpurun_out/timeline_a6_16.txt
