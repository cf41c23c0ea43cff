#!/bin/bash
mkdir -p gpurun_out/gbgc5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gbgc_chain_kernel -c 1 -o gpurun_out/gbgc5/gbgc_chain python tools/gbgc_smoke.py 148 10 30 > gpurun_out/gbgc5/ncu.log 2>&1
tail -3 gpurun_out/gbgc5/ncu.log
ls -la gpurun_out/gbgc5
