#!/bin/bash
# two GPUs: sharded-product tests of both programs (one shard per GPU), cfg4 weak-scaling line under torchrun, the ST binary on 16384 elements
mkdir -p gpurun_out/r2n
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gbgc_dropin.py tests/test_gpu_posterior.py -m gpu -q -W ignore -k sharded 2>&1 | tail -3 | tee gpurun_out/r2n/pytest_sharded.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2n/bench_cfg4_2gpu.json 2> gpurun_out/r2n/bench_cfg4_2gpu.err
tail -c 700 gpurun_out/r2n/bench_cfg4_2gpu.json; tail -2 gpurun_out/r2n/bench_cfg4_2gpu.err
python - <<'PY'
import os, sys, time, subprocess, shutil
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from phyloacc_b200 import synth
tree = synth.make_tree()
batch = synth.simulate(tree, C=8192, length="cfg2", seed=77)
out = {}
for tag, env in (("2gpu", {}), ("1gpu", {"PA_DEVICES": "1"})):
    d = f"/tmp/pa_prod_{tag}"
    shutil.rmtree(d, ignore_errors=True)
    cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=16))
    t0 = time.time()
    r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, env=dict(os.environ, PA_TIMING="1", **env))
    out[tag] = time.time() - t0
    print(tag, "rc", r.returncode, "wall %.1f s" % out[tag], [l for l in r.stdout.split("\n") if "shards" in l], r.stderr[-600:])
import filecmp
same = all(filecmp.cmp(f"/tmp/pa_prod_2gpu/out/{f}", f"/tmp/pa_prod_1gpu/out/{f}", shallow=False) for f in os.listdir("/tmp/pa_prod_1gpu/out"))
open("gpurun_out/r2n/product_2gpu.txt", "w").write(f"PhyloAcc-ST-b200 on 8192 cfg2 elements, 1500 iterations x 3 models: 2 GPUs {out['2gpu']:.1f} s, 1 GPU {out['1gpu']:.1f} s wall, byte-identical files: {same}\n")
PY
cat gpurun_out/r2n/product_2gpu.txt
