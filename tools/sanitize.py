#!/usr/bin/env python
"""Tiny chain + replay run per compiled shared-memory layout, sized for compute-sanitizer (racecheck / memcheck / synccheck):

   compute-sanitizer --tool racecheck --print-limit 50 python tools/sanitize.py 43
   compute-sanitizer --tool memcheck  python tools/sanitize.py 8 43 70 110 230

Each taxon count selects one layout of chain_kernel (pa_chain.cu: pick_variant): 8 -> <32>, 43 -> <96>, 70 -> <160>, 110 -> <240>,
230 -> <480>.  The elements carry gaps, missing species and IUPAC ambiguity codes so that the rare paths run too."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phyloacc_b200 as pab  # noqa: E402
from phyloacc_b200 import synth  # noqa: E402

taxa = [int(x) for x in sys.argv[1:]] or [8, 43]
iters = int(os.environ.get("PA_SAN_ITERS", "12"))
elems = int(os.environ.get("PA_SAN_ELEMS", "3"))
length = int(os.environ.get("PA_SAN_LEN", "70"))
for S in taxa:
    tree = synth.make_tree(S=S, seed=9, n_out=max(2, S // 6), n_target=max(2, S // 5))
    batch = synth.simulate(tree, C=elems, length=length, seed=5, gap_frac=0.03, missing_frac=0.05, ambig_frac=0.01)
    prm = pab.default_params()
    prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed, prm.num_thread = iters // 3, iters - iters // 3, 4, 1, 2
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                            conserve=tree.conserve, device=0)
    eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    ids = np.arange(batch.C, dtype=np.int32)
    res = eng.run(ids)
    ok = np.isfinite(res.loglik_max[res.filtered == 0]).all()
    print(f"taxa={S} elems={elems} iters={iters} finite={bool(ok)} launches={eng.counters()['kernel_launches']}", flush=True)
    eng.close()
