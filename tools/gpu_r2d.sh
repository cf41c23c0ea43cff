mkdir -p gpurun_out/r2d
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2d/pytest_gpu.txt
cat gpurun_out/r2d/pytest_gpu.txt
PA_TIMING=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
tail -5 gpurun_out/r2d/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d/bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['frac'])
PY
