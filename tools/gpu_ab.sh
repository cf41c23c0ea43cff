# usage: gpu_ab.sh lib1.so lib2.so ...   (A/B timing of chain-kernel builds: profile_chain on 200 bp and on cfg2 lengths)
mkdir -p gpurun_out/ab
for lib in "$@"; do
  echo "=== $lib"
  PA_LIB=$PWD/phyloacc_b200/$lib python tools/profile_chain.py 2>&1 | tail -2
  PA_LIB=$PWD/phyloacc_b200/$lib PA_PROF_LEN=cfg2 PA_PROF_ELEMS=1200 python tools/profile_chain.py 2>&1 | tail -2
done
