#!/bin/bash
mkdir -p gpurun_out/gbgc4
for n in 592 1184 2048; do
timeout 600 python tools/gbgc_smoke.py $n 50 150 2>&1 | grep "chains\|cycles\|rror" | tee -a gpurun_out/gbgc4/smoke.txt
done
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/gbgc4/bench_cfg4.json 2> gpurun_out/gbgc4/bench_cfg4.err
tail -c 3000 gpurun_out/gbgc4/bench_cfg4.json
tail -5 gpurun_out/gbgc4/bench_cfg4.err
