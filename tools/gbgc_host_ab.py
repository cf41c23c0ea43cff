"""A/B of two builds of the library on the same GBGC chains: every summary must be bit-identical (the chains are keyed by seed and element),
wall time of pa_engine_gbgc_run tells what the host side costs.  usage: PA_LIB=<lib> python tools/gbgc_host_ab.py <out.npz>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phyloacc_b200 as pab
from phyloacc_b200 import synth
import tempfile
tree = synth.make_tree()
batch = synth.simulate(tree, C=1024, length=200, seed=5, gap_frac=0.01, missing_frac=0.03, ambig_frac=0.002, gbgc={})
d = tempfile.mkdtemp()
cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=5, MCMC=10, CHAIN=1, GAPCHAR="-"))
eng = pab.Engine.from_config(cfg, device=0)
eng.gbgc_run(list(range(64)), tree.targets, num_burn=20, num_mcmc=40, seed=3)  # warm-up
t = time.time()
out = eng.gbgc_run(list(range(1024)), tree.targets, num_burn=200, num_mcmc=800, seed=3)
dt = time.time() - t
print(f"{os.environ.get('PA_LIB', 'default')}: wall {dt:.2f} s, kernel {out['kernel_ms'] / 1e3:.2f} s")
np.savez(sys.argv[1], **{k: out[k] for k in ("loglik", "loglik_max", "med", "post_b", "post_z", "max_z", "max_b", "accept")})
