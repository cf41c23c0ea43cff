#!/bin/bash
mkdir -p gpurun_out/final2
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -4 | tee gpurun_out/final2/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final2/smoke.txt
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/final2/bench_cfg4.json 2> gpurun_out/final2/bench_cfg4.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/final2/bench_cfg4.json").read().strip().split("\n")[-1])
print("cfg4 value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "cpu", l["cpu_baseline"]["value"], "launches", l["gpu_launches"])
print(l["roofline"]["phase_kcycles_per_chain_iteration"])
PY
