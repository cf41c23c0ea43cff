set -x
mkdir -p gpurun_out/r2c
for k in 1 2 3; do PA_CTAS_PER_SM_RUNTIME=$k PA_PROF_ELEMS=$((148*k/3+1)) python tools/profile_chain.py > gpurun_out/r2c/occ_$k.txt 2>&1; done
cat gpurun_out/r2c/occ_*.txt
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/r2c/chain python tools/profile_chain.py > gpurun_out/r2c/ncu.log 2>&1
tail -3 gpurun_out/r2c/ncu.log
