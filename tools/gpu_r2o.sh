#!/bin/bash
mkdir -p gpurun_out/r2o
python - <<'PY'
import os, sys, time, subprocess, shutil
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from phyloacc_b200 import synth
tree = synth.make_tree()
for C in (50, 2048):
    batch = synth.simulate(tree, C=C, length="cfg2", seed=77)
    d = f"/tmp/pa_prod_{C}"
    shutil.rmtree(d, ignore_errors=True)
    cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=16))
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, env=dict(os.environ, PA_TIMING="1", PA_DEVICES="1"))
        print(C, "elements, run", rep, "rc", r.returncode, "wall %.2f s" % (time.time() - t0))
        print(r.stderr[-1500:])
PY
