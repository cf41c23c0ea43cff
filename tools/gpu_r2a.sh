set -x
mkdir -p gpurun_out/r2a
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_base.json 2> gpurun_out/r2a/bench_base.err
for S in 8 43 70 110 230; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize.py $S > gpurun_out/r2a/racecheck_$S.txt 2>&1
  timeout 300 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize.py $S > gpurun_out/r2a/memcheck_$S.txt 2>&1
done
timeout 300 compute-sanitizer --tool synccheck --print-limit 30 python tools/sanitize.py 43 > gpurun_out/r2a/synccheck_43.txt 2>&1
tail -3 gpurun_out/r2a/*.txt
