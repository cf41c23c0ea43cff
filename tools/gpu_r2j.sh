# 8-GPU box: final weak-scaling line with the pipelined e2e and the whole 280k-element cfg2 job; the multi-GPU product test
mkdir -p gpurun_out/r2j
python -m pytest tests/test_gpu_summaries.py tests/test_gpu_posterior.py -m gpu -q -W ignore 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --full-run > gpurun_out/r2j/bench_8gpu_full.json 2> gpurun_out/r2j/bench_8gpu_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j/bench_8gpu_full.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'], d['full_run'], d['cpu_baseline']['value'])
PY
