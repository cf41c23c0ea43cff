#!/bin/bash
mkdir -p gpurun_out/gbgc8
timeout 900 python -m pytest tests/test_gbgc_dropin.py -m gpu -x -q -W ignore 2>&1 | tail -15 | tee gpurun_out/gbgc8/pytest_dropin.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/gbgc_smoke.py 3 3 6" >> gpurun_out/gbgc8/sanitizer_gbgc.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/gbgc_smoke.py 3 3 6 2>&1 | grep -v "^Loading\|^$" | tail -12 >> gpurun_out/gbgc8/sanitizer_gbgc.txt
done
tail -40 gpurun_out/gbgc8/sanitizer_gbgc.txt
