#!/bin/bash
mkdir -p gpurun_out/gbgc2
timeout 600 python tools/gbgc_smoke.py 592 50 150 2>&1 | grep "chains\|cycles" | tee gpurun_out/gbgc2/smoke592.txt
timeout 600 python -m pytest tests/test_gbgc_chain.py -m gpu -x -q -W ignore -k "draw_for_draw or replay" 2>&1 | tail -3
