#!/usr/bin/env python
"""Wall-clock split of the end-to-end path on the bench workload: front-end + upload, chains, trace download + Eval2."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import phyloacc_b200 as pab  # noqa: E402

elems = int(os.environ.get("PA_ELEMS", "2048"))
tree, batch = bench.make_batch(elems, seed=1000)
prm = pab.default_params()
prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed = bench.BURNIN, bench.MCMC, 500, 1
prm.num_thread = os.cpu_count() or 1
eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup, conserve=tree.conserve, device=0)
ids = np.arange(batch.C, dtype=np.int32)
for rep in range(int(os.environ.get("PA_REPS", "4"))):
    t0 = time.time()
    eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    t1 = time.time()
    eng.sample(ids)
    torch.cuda.synchronize()
    t2 = time.time()
    res = eng.collect()
    t3 = time.time()
    print(f"rep {rep}: load_alignment {t1 - t0:.3f}s  chains {t2 - t1:.3f}s  collect (D2H + Eval2 + summaries) {t3 - t2:.3f}s  threads={prm.num_thread}")
