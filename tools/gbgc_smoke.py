"""Tiny GBGC chain run (for compute-sanitizer and quick timing): python tools/gbgc_smoke.py [n_elements] [burn] [mcmc]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import phyloacc_b200 as pab
from phyloacc_b200 import synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mcmc = int(sys.argv[3]) if len(sys.argv) > 3 else 10
tree = synth.make_tree()
batch = synth.simulate(tree, C=C, length=200, seed=5, gap_frac=0.01, missing_frac=0.03, ambig_frac=0.002)
eng = pab.Engine.from_batch(batch, device=0) if hasattr(pab.Engine, "from_batch") else None
if eng is None:
    import tempfile

    d = tempfile.mkdtemp()
    cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=burn, MCMC=mcmc, CHAIN=1, GAPCHAR="-"))
    eng = pab.Engine.from_config(cfg, device=0)
t = time.time()
out = eng.gbgc_run(list(range(C)), tree.targets, num_burn=burn, num_mcmc=mcmc, seed=3)
dt = time.time() - t
its = out["chain_iterations"]
print(f"chains {3 * C} iterations {its} kernel {out['kernel_ms']:.1f} ms  wall {dt:.2f} s  -> {its / 3 / (out['kernel_ms'] / 1e3):.0f} element-iterations/s (kernel)")
names = "setup monitor down init_emis init_msg build emis msg draw up elbo commit".split()
print("cycles per chain-iteration:", ", ".join(f"{n} {c / max(its, 1) / 1e3:.1f} k" for n, c in zip(names, out["phase_cycles"])), "| total %.1f k" % (sum(out["phase_cycles"]) / max(its, 1) / 1e3))
print("loglik", out["loglik"][:2], "med", out["med"][0], "accept", out["accept"][0])
