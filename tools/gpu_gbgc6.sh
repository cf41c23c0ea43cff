#!/bin/bash
mkdir -p gpurun_out/gbgc6
for lib in libphyloacc_b200.so libpa_g4.so; do
echo "== $lib" | tee -a gpurun_out/gbgc6/smoke.txt
PA_LIB=$PWD/phyloacc_b200/$lib timeout 600 python tools/gbgc_smoke.py 1184 50 150 2>&1 | grep "chains\|cycles\|rror" | tee -a gpurun_out/gbgc6/smoke.txt
done
timeout 900 python -m pytest tests/test_gbgc_chain.py tests/test_gbgc_oracle.py -m gpu -x -q -W ignore 2>&1 | tail -15 | tee gpurun_out/gbgc6/pytest.txt
