#!/bin/bash
# final single-GPU lines of round 2: default workload (cfg2) and cfg4, both with their cpu_baseline; reference arm of cfg4
mkdir -p gpurun_out/r2m
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m/bench_cfg2.json 2> gpurun_out/r2m/bench_cfg2.err
tail -c 600 gpurun_out/r2m/bench_cfg2.json; tail -2 gpurun_out/r2m/bench_cfg2.err
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r2m/bench_cfg4.json 2> gpurun_out/r2m/bench_cfg4.err
tail -c 900 gpurun_out/r2m/bench_cfg4.json; tail -2 gpurun_out/r2m/bench_cfg4.err
timeout 600 python bench.py --impl reference --workload cfg4 --steps 3 --warmup 1 > gpurun_out/r2m/bench_ref_cfg4.json 2> gpurun_out/r2m/bench_ref_cfg4.err
head -c 700 gpurun_out/r2m/bench_ref_cfg4.json
