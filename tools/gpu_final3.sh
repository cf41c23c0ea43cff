#!/bin/bash
mkdir -p gpurun_out/final3
timeout 900 python -m pytest tests/test_gbgc_chain.py tests/test_gbgc_oracle.py tests/test_gbgc_dropin.py -m gpu -x -q -W ignore 2>&1 | tail -3 | tee gpurun_out/final3/pytest_gbgc.txt
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/final3/bench_cfg4.json 2> gpurun_out/final3/bench_cfg4.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/final3/bench_cfg4.json").read().strip().split("\n")[-1])
print("cfg4 value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "cpu", l["cpu_baseline"]["value"], "launches", l["gpu_launches"])
PY
