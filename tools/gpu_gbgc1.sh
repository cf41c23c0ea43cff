#!/bin/bash
# GBGC chain: parity tests on the GPU (draw for draw against the numpy oracle, replay invariant, reference seeds)
mkdir -p gpurun_out/gbgc1
timeout 900 python -m pytest tests/test_gbgc_chain.py -m gpu -x -q -W ignore 2>&1 | tail -30 | tee gpurun_out/gbgc1/pytest.txt
if grep -q "illegal\|failed" gpurun_out/gbgc1/pytest.txt; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/gbgc_smoke.py 2 2 3 2>&1 | grep -v "^Loading\|^#" | head -60 | tee gpurun_out/gbgc1/memcheck.txt
fi
timeout 300 python tools/gbgc_smoke.py 592 50 150 2>&1 | tail -3 | tee gpurun_out/gbgc1/smoke.txt
