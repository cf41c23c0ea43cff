# usage: gpu_ab2.sh lib1.so lib2.so ...  (A/B of chain-kernel builds through PA_LIB: cycles per chain-iteration by phase, 200 bp and cfg2 lengths)
for lib in "$@"; do
  echo "=== $lib"
  PA_LIB=$PWD/phyloacc_b200/$lib PA_PROF_ELEMS=592 python tools/profile_chain.py 2>&1 | tail -2
  PA_LIB=$PWD/phyloacc_b200/$lib PA_PROF_LEN=cfg2 PA_PROF_ELEMS=2048 PA_PROF_ITERS=200 python tools/profile_chain.py 2>&1 | tail -2
done
