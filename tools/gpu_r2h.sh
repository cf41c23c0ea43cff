mkdir -p gpurun_out/r2h
python -m pytest tests -m gpu -q -W ignore 2>&1 | tail -4
PA_PROF_ELEMS=592 python tools/profile_chain.py 2>&1 | tail -2
python bench.py --steps 6 --warmup 3 > gpurun_out/r2h/bench_cfg2.json 2> gpurun_out/r2h/bench_cfg2.err
tail -3 gpurun_out/r2h/bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h/bench_cfg2.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
