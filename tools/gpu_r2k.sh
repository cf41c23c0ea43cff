#!/bin/bash
# round-2 late evidence: whole GPU suite, cfg4 bench line, launch list and full ncu capture of the GBGC chain kernel
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -8 | tee gpurun_out/r2k/pytest_gpu.txt
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r2k/bench_cfg4.json 2> gpurun_out/r2k/bench_cfg4.err
tail -c 2500 gpurun_out/r2k/bench_cfg4.json; tail -3 gpurun_out/r2k/bench_cfg4.err
PA_BENCH_ELEMS=592 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2k/launches_cfg4.csv python bench.py --workload cfg4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2k/bench_cfg4_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gbgc_chain_kernel -c 1 -o gpurun_out/r2k/gbgc_chain python tools/gbgc_smoke.py 148 10 30 > gpurun_out/r2k/ncu_gbgc.log 2>&1
ls -la gpurun_out/r2k
