mkdir -p gpurun_out/r2i
cd /root/repo
python -m pytest tests -m gpu -q -W ignore 2>&1 | tail -5 > gpurun_out/r2i/pytest_gpu.txt
cat gpurun_out/r2i/pytest_gpu.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r2i/clocks.csv &
SMI=$!
python bench.py --steps 5 --warmup 3 > gpurun_out/r2i/bench_cfg2.json 2> gpurun_out/r2i/bench_cfg2.err
python bench.py --steps 5 --warmup 3 --workload cfg1 > gpurun_out/r2i/bench_cfg1.json 2> gpurun_out/r2i/bench_cfg1.err
python bench.py --steps 3 --warmup 3 --workload cfg3 > gpurun_out/r2i/bench_cfg3.json 2> gpurun_out/r2i/bench_cfg3.err
kill $SMI
python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/r2i/bench_ref_cfg2.json 2> gpurun_out/r2i/bench_ref_cfg2.err
for f in gpurun_out/r2i/bench_*.json; do echo $f; head -c 300 $f; echo; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2i/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2i/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/r2i/chain python tools/profile_chain.py > gpurun_out/r2i/ncu_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:summary_kernel -c 1 -o gpurun_out/r2i/summary python tools/profile_chain.py > gpurun_out/r2i/ncu_summary.log 2>&1
# DRAM traffic of one bench-size launch (2048 cfg2 elements x 3 x 1500 iterations)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:chain_kernel -c 1 --csv --log-file gpurun_out/r2i/traffic_cfg2.csv python tools/profile_chain.py > /dev/null 2>&1 <<< ""
PA_PROF_LEN=cfg2 PA_PROF_ELEMS=2048 PA_PROF_ITERS=1500 PA_PROF_REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:chain_kernel -c 1 --csv --log-file gpurun_out/r2i/traffic_cfg2.csv python tools/profile_chain.py > gpurun_out/r2i/traffic_cfg2.log 2>&1
for S in 8 43; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize.py $S > gpurun_out/r2i/racecheck_$S.txt 2>&1
  timeout 300 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize.py $S > gpurun_out/r2i/memcheck_$S.txt 2>&1
done
for S in 70 110 230; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize.py $S > gpurun_out/r2i/racecheck_$S.txt 2>&1
done
grep -h "SUMMARY" gpurun_out/r2i/*check_*.txt
