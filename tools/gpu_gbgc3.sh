#!/bin/bash
mkdir -p gpurun_out/gbgc14
timeout 600 python tools/gbgc_smoke.py 1184 50 150 2>&1 | grep "chains\|cycles\|rror" | tee gpurun_out/gbgc14/smoke592.txt
timeout 900 python -m pytest tests/test_gbgc_chain.py tests/test_gbgc_oracle.py -m gpu -x -q -W ignore 2>&1 | tail -15 | tee gpurun_out/gbgc14/pytest.txt
