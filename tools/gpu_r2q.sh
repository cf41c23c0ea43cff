#!/bin/bash
# several GPUs: the drop-in binary on 8192 cfg2 elements (one shard per GPU) with the stage times
mkdir -p gpurun_out/r2q
nvidia-smi -L | wc -l
python - <<'PY'
import os, sys, time, subprocess, shutil
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from phyloacc_b200 import synth
tree = synth.make_tree()
batch = synth.simulate(tree, C=8192, length="cfg2", seed=77)
d = "/tmp/pa_prod_8gpu"
shutil.rmtree(d, ignore_errors=True)
cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=32))
res = []
for rep in range(2):
    shutil.rmtree(d + "/out", ignore_errors=True); os.makedirs(d + "/out")
    t0 = time.time()
    r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, env=dict(os.environ, PA_TIMING="1"))
    dt = time.time() - t0
    res.append(dt)
    print("run", rep, "rc", r.returncode, "wall %.1f s" % dt, [l for l in r.stdout.split("\n") if "shards" in l])
    print("\n".join(l for l in r.stderr.split("\n") if "timing" in l and "collect" not in l))
open("gpurun_out/r2q/product_8gpu.txt", "w").write(f"PhyloAcc-ST-b200 on 8192 cfg2 elements, 1500 iterations x 3 models, N GPUs (one shard each), NUM_THREAD 32: {res[0]:.1f} s and {res[1]:.1f} s wall (one GPU: 10.3 s, two: 8.2 s)\n")
PY
cat gpurun_out/r2q/product_8gpu.txt
