#!/bin/bash
mkdir -p gpurun_out/r2r
python - <<'PY'
import os, sys, time, subprocess, shutil
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from phyloacc_b200 import synth
tree = synth.make_tree()
batch = synth.simulate(tree, C=4096, length="cfg2", seed=77)
d = "/tmp/pa_prod_x"
shutil.rmtree(d, ignore_errors=True)
cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=16))
for tag, env in (("1 shard", {"PA_DEVICES": "1"}), ("4 shards on GPU 0", {"PA_DEVICE_LIST": "0,0,0,0"}), ("1 shard again", {"PA_DEVICES": "1"})):
    t0 = time.time()
    r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, env=dict(os.environ, PA_TIMING="1", **env))
    print("==", tag, "rc", r.returncode, "wall %.1f s" % (time.time() - t0))
    print("\n".join(l for l in r.stderr.split("\n") if "timing" in l and "device summaries" not in l))
PY
