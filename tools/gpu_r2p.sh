#!/bin/bash
# end-of-round check: whole GPU suite, smoke(), cfg4 bench line with its cpu_baseline, ncu summary capture of the final GBGC kernel
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -4 | tee gpurun_out/r2p/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2p/smoke.txt
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r2p/bench_cfg4.json 2> gpurun_out/r2p/bench_cfg4.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2p/bench_cfg4.json").read().strip().split("\n")[-1])
print("cfg4 value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "cpu", l["cpu_baseline"]["value"])
print(l["roofline"]["phase_kcycles_per_chain_iteration"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gbgc_chain_kernel -c 1 -o gpurun_out/r2p/gbgc_chain python tools/gbgc_smoke.py 148 10 30 > gpurun_out/r2p/ncu_gbgc.log 2>&1
ls -la gpurun_out/r2p | tail -8
