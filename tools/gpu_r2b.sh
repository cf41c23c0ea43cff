set -x
mkdir -p gpurun_out/r2b
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b/pytest_gpu.txt
cat gpurun_out/r2b/pytest_gpu.txt
python tools/profile_chain.py > gpurun_out/r2b/profile_200.txt 2>&1
PA_PROF_LEN=cfg2 PA_PROF_ELEMS=600 python tools/profile_chain.py > gpurun_out/r2b/profile_cfg2.txt 2>&1
cat gpurun_out/r2b/profile_*.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
head -c 400 gpurun_out/r2b/bench.json
