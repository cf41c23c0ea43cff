python -m pytest tests/test_gpu_chain.py tests/test_gpu_summaries.py tests/test_gpu_invariants.py -m gpu -q -W ignore 2>&1 | tail -3
for lib in libpa_deep1.so libphyloacc_b200.so; do
  echo "=== $lib"
  PA_LIB=$PWD/phyloacc_b200/$lib PA_TAXA=240 PA_ELEMS=592 PA_ITERS=200 python tools/deep_tree_rate.py 2>&1 | tail -1
  PA_LIB=$PWD/phyloacc_b200/$lib PA_TAXA=200 PA_ELEMS=592 PA_ITERS=200 python tools/deep_tree_rate.py 2>&1 | tail -1
done
