#!/bin/bash
mkdir -p gpurun_out/gbgc9
for lib in libphyloacc_b200.so libpa_g192.so libpa_g256.so; do
echo "== $lib" | tee -a gpurun_out/gbgc9/smoke.txt
PA_LIB=$PWD/phyloacc_b200/$lib timeout 600 python tools/gbgc_smoke.py 1184 50 150 2>&1 | grep "chains\|cycles\|rror" | tee -a gpurun_out/gbgc9/smoke.txt
done
