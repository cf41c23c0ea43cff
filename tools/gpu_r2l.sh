#!/bin/bash
# fast writers: byte-level output tests, then the drop-in binary on 16384 cfg2 elements (1 GPU), wall time by stage from NVTX-free timestamps
mkdir -p gpurun_out/r2l
timeout 1200 python -m pytest tests/test_gpu_outputs.py tests/test_gpu_posterior.py tests/test_gbgc_dropin.py tests/test_hyper.py -m gpu -x -q -W ignore 2>&1 | tail -5 | tee gpurun_out/r2l/pytest.txt
python - <<'PY'
import os, sys, time, subprocess, shutil
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from phyloacc_b200 import synth
tree = synth.make_tree()
batch = synth.simulate(tree, C=16384, length="cfg2", seed=77)
d = "/tmp/pa_prod_1gpu"
shutil.rmtree(d, ignore_errors=True)
cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=16))
t0 = time.time()
r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, env=dict(os.environ, PA_DEVICES="1"))
dt = time.time() - t0
print("rc", r.returncode, "wall %.1f s" % dt)
open("gpurun_out/r2l/product_1gpu.txt", "w").write(f"PhyloAcc-ST-b200 on 16384 cfg2 elements, 1500 iterations x 3 models, 1 GPU, NUM_THREAD 16: {dt:.1f} s wall (round-2 v1 with iostream writers: 24.1 s)\n")
PY
cat gpurun_out/r2l/product_1gpu.txt
