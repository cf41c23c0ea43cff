#!/usr/bin/env python
"""Fixture for the PhyloAcc-GT groundwork: oracle/_ref/gt_probe (the UNMODIFIED sources of src/PhyloAcc-GT, built by
`make -C oracle ref-gt`) on 4 elements of 12 taxa x 120 bp (gaps, missing species, ambiguity codes) at 6 states each: gene trees with the
species topology (GTree::initTree_Sptop) and draws from the multispecies coalescent (GTree::initTree, theta = 0.01), random Z and rates.
Per state the probe prints the embedded gene tree, the coalescent prior (GTree::priorTree), log P(missing | Z) (get_logpZ1) and the
log-likelihood of the upward pass (GTree::Update_Tg) as hex floats.
Run here:  python tests/golden/make_gt_golden.py  ->  tests/golden/gt12/"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from phyloacc_b200 import synth  # noqa: E402

PROBE = os.path.join(ROOT, "oracle", "_ref", "gt_probe")
OUT = os.path.join(ROOT, "tests", "golden", "gt12")
THETA = 0.01


def main():
    tree = synth.make_tree(S=12, seed=5, n_out=2, n_target=3)
    batch = synth.simulate(tree, C=4, length=120, seed=11, gap_frac=0.03, missing_frac=0.08, ambig_frac=0.01)
    work = tempfile.mkdtemp(prefix="pa_gt_")
    cfg = synth.write_files(batch, work, prefix="sim", params=dict(BURNIN=15, MCMC=30, CHAIN=1, SEED=2, NUM_THREAD=1, GAPCHAR="-"))
    # species tree in coalescent units = substitution lengths / (theta / 2) (PhyloAcc-GT main.cpp:376-400)
    coal = re.sub(r":([0-9.]+)", lambda m: ":%.6f" % (float(m.group(1)) / (THETA / 2)), tree.newick)
    with open(os.path.join(work, "sim_coal.nwk"), "w") as f:
        f.write(coal + "\n")
    with open(cfg, "a") as f:
        f.write(f"TREE_IN_COALESCENT_UNIT {work}/sim_coal.nwk\nTHIN 10\n")
    N = 2 * tree.S - 1
    rng = np.random.default_rng(1)
    lines = []
    for c in range(batch.C):
        lines.append(f"element {c}")
        for k in range(6):
            Z = rng.integers(0, 3, N)
            lines.append(f"state {(1, 2, 0)[k % 3]} {0 if k < 2 else 1} {0.5 + rng.random() * 2:.6f} {0.1 + rng.random() * 0.8:.6f} " + " ".join(map(str, Z)))
    script = os.path.join(work, "script.txt")
    with open(script, "w") as f:
        f.write("\n".join(lines) + "\n")
    r = subprocess.run([PROBE, cfg, script], capture_output=True, text=True, check=True)
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for fn in ("sim.mod", "sim.fa", "sim.bed", "sim.id", "sim_coal.nwk"):
        shutil.copy(os.path.join(work, fn), os.path.join(OUT, fn))
    with open(cfg) as f:
        keep = [l for l in f.read().split("\n") if l and not l.split()[0].endswith("_FILE") and l.split()[0] not in ("RESULT_FOLDER", "TREE_IN_COALESCENT_UNIT")]
    with open(os.path.join(OUT, "params.cfg"), "w") as f:
        f.write("\n".join(keep) + "\n")
    with open(os.path.join(OUT, "script.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(OUT, "probe.txt"), "w") as f:
        f.write(r.stdout)
    # a second, separate run: two states per element (species topology, coalescent draw) with one sampled substitution history each and the
    # Z emissions BPP_C::getEmission derives from it ("hist")
    rng = np.random.default_rng(2)
    hlines = []
    for c in range(batch.C):
        hlines.append(f"element {c}")
        for k in range(2):
            Z = rng.integers(0, 3, N)
            hlines.append(f"state 1 {k} {0.5 + rng.random() * 2:.6f} {0.1 + rng.random() * 0.8:.6f} " + " ".join(map(str, Z)) + " hist")
    hscript = os.path.join(work, "script_hist.txt")
    with open(hscript, "w") as f:
        f.write("\n".join(hlines) + "\n")
    rh = subprocess.run([PROBE, cfg, hscript], capture_output=True, text=True, check=True)
    with open(os.path.join(OUT, "script_hist.txt"), "w") as f:
        f.write("\n".join(hlines) + "\n")
    with open(os.path.join(OUT, "probe_hist.txt"), "w") as f:
        f.write(rh.stdout)
    print("wrote", OUT, len(r.stdout), "bytes of probe output")


if __name__ == "__main__":
    main()
