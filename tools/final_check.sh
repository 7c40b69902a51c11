#!/bin/bash
# What the driver runs at round end, in one go: smoke(), the GPU test suite, the default bench line.
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
mkdir -p gpurun_out
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"],
      d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
