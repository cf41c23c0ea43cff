#!/bin/bash
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -4 | tee gpurun_out/final/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final/smoke.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_cfg2.json 2> gpurun_out/final/bench_cfg2.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/final/bench_cfg2.json").read().strip().split("\n")[-1])
print("cfg2 value", l["value"], "e2e", l["e2e"]["value"], "frac", l["roofline"]["frac"])
PY
