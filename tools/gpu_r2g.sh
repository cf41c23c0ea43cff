# 8-GPU box: sharded product test, weak-scaling bench lines with the whole 280k-element cfg2 job through the public API
mkdir -p gpurun_out/r2g
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_posterior.py -m gpu -q -W ignore -k sharded 2>&1 | tail -3 > gpurun_out/r2g/pytest_sharded.txt
cat gpurun_out/r2g/pytest_sharded.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --full-run > gpurun_out/r2g/bench_8gpu_full.json 2> gpurun_out/r2g/bench_8gpu_full.err
tail -c 1500 gpurun_out/r2g/bench_8gpu_full.json
tail -3 gpurun_out/r2g/bench_8gpu_full.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 2 > gpurun_out/r2g/bench_8gpu_reference.json 2> gpurun_out/r2g/bench_8gpu_reference.err
head -c 600 gpurun_out/r2g/bench_8gpu_reference.json
# the drop-in binary on a genome-shaped id file: 16384 elements over 8 GPUs vs 1 GPU, byte-identical files
python - <<'PY'
import os, sys, time, subprocess, filecmp, shutil
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from phyloacc_b200 import synth
tree = synth.make_tree()
batch = synth.simulate(tree, C=16384, length="cfg2", seed=77)
out = {}
for tag, env in (("8gpu", {}), ("1gpu", {"PA_DEVICES": "1"})):
    d = f"/tmp/pa_prod_{tag}"
    shutil.rmtree(d, ignore_errors=True)
    cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=32))
    t0 = time.time()
    r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, env=dict(os.environ, **env))
    out[tag] = time.time() - t0
    print(tag, "rc", r.returncode, "wall %.1f s" % out[tag], [l for l in r.stdout.split("\n") if "shards" in l])
same = all(filecmp.cmp(f"/tmp/pa_prod_8gpu/out/{f}", f"/tmp/pa_prod_1gpu/out/{f}", shallow=False) for f in os.listdir("/tmp/pa_prod_1gpu/out"))
print("byte-identical:", same, " speed-up of the whole program (parse + front-end + chains + files): %.2fx" % (out["1gpu"] / out["8gpu"]))
open("gpurun_out/r2g/product_8gpu.txt", "w").write(f"PhyloAcc-ST-b200 on 16384 cfg2 elements, 1500 iterations x 3 models: 8 GPUs {out['8gpu']:.1f} s, 1 GPU {out['1gpu']:.1f} s, byte-identical files: {same}\n")
PY
