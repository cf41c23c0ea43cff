#!/usr/bin/env python
"""Throughput of the deep-tree layout (BASELINE config 3 shape: 240 taxa, 200 bp) -- a measured side number, not the bench line."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phyloacc_b200 as pab  # noqa: E402
from phyloacc_b200 import synth  # noqa: E402

S = int(os.environ.get("PA_TAXA", "240"))
elems = int(os.environ.get("PA_ELEMS", "296"))
iters = int(os.environ.get("PA_ITERS", "100"))
tree = synth.make_tree(S=S, seed=9, n_out=max(2, S // 6), n_target=max(2, S // 5))
batch = synth.simulate(tree, C=elems, length=200, seed=5)
prm = pab.default_params()
prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed, prm.num_thread = iters // 2, iters - iters // 2, 500, 1, 8
eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup, conserve=tree.conserve, device=0)
eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
ids = np.arange(batch.C, dtype=np.int32)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    eng.sample(ids)
    e1.record(stream)
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
eng.collect()
c = eng.counters()
fma = 36.0 * c["node_pattern_updates"] + 4.0 * c["root_evals"] + 80.0 * c["pt_builds"] + 4.0 * c["history_draws"]
print(f"taxa={S} elems={elems} iters={iters} kernel={dt:.3f}s el-iter/s={elems * 3 * iters / dt:.0f} algorithmic TFLOP/s={2 * fma / dt / 1e12:.3f} launches={c['kernel_launches']}")
