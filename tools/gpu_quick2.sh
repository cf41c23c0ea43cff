python -m pytest tests/test_gpu_summaries.py tests/test_gpu_chain.py tests/test_gpu_outputs.py -m gpu -q -W ignore 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:summary_kernel -c 3 --csv --log-file gpurun_out/summary_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep summary_kernel gpurun_out/summary_launches.csv | awk -F'","' '{print $5, $NF}'
