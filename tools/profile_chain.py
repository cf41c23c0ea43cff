#!/usr/bin/env python
"""Small, fixed workload for ncu captures of the chain kernel (one launch, ~0.1-0.3 s):
   ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/chain python tools/profile_chain.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phyloacc_b200 as pab  # noqa: E402
from phyloacc_b200 import synth  # noqa: E402

elems = int(os.environ.get("PA_PROF_ELEMS", "200"))
iters = int(os.environ.get("PA_PROF_ITERS", "100"))
length = os.environ.get("PA_PROF_LEN", "200")
length = "cfg2" if length == "cfg2" else int(length)
tree = synth.make_tree()
batch = synth.simulate(tree, C=elems, length=length, seed=77)
prm = pab.default_params()
prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed = iters // 2, iters - iters // 2, 500, 1
prm.num_thread = 4
eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup, conserve=tree.conserve, device=0)
eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
ids = np.arange(batch.C, dtype=np.int32)
import torch  # noqa: E402  (stream / events only)

reps = int(os.environ.get("PA_PROF_REPS", "3"))
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
for _ in range(reps):  # the last repetition is the timed one (the first pays module load and allocation)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    eng.sample(ids)
    e1.record(stream)
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
res = eng.collect()
c = eng.counters()
print(f"elems={elems} iters={iters} kernel={dt:.4f}s  el-iter/s={elems * 3 * iters / dt:.0f}  counters={c}")
ph = eng.phase_cycles()
tot = sum(v for k, v in ph.items() if k != 'c_pass_only')
print("phase cycles per chain-iteration:", {k: round(v / max(c["chain_iterations"], 1)) for k, v in ph.items()}, "total", round(tot / max(c["chain_iterations"], 1)))
