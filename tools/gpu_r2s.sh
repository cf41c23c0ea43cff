#!/bin/bash
mkdir -p gpurun_out/r2s
for lib in libpa_old.so libphyloacc_b200.so libpa_old.so libphyloacc_b200.so; do
  echo "=== $lib" | tee -a gpurun_out/r2s/ab.txt
  PA_LIB=$PWD/phyloacc_b200/$lib python tools/profile_chain.py 2>&1 | tail -2 | tee -a gpurun_out/r2s/ab.txt
done
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_invariants.py tests/test_gpu_summaries.py tests/test_gpu_replay.py -m gpu -x -q -W ignore 2>&1 | tail -4 | tee gpurun_out/r2s/pytest.txt
