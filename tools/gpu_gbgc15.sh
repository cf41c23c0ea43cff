#!/bin/bash
mkdir -p gpurun_out/gbgc15
for lib in libpa_old.so libphyloacc_b200.so libpa_old.so libphyloacc_b200.so; do
echo "== $lib" | tee -a gpurun_out/gbgc15/smoke.txt
PA_LIB=$PWD/phyloacc_b200/$lib timeout 600 python tools/gbgc_smoke.py 1184 50 150 2>&1 | grep "chains\|cycles\|rror" | tee -a gpurun_out/gbgc15/smoke.txt
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
