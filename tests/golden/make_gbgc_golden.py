#!/usr/bin/env python
"""Golden vectors for the GC-biased model's per-branch matrices (PhyloAcc-ST-GBGC), from the UNMODIFIED reference:
oracle/_ref/gbgc_probe (make -C oracle ref-gbgc) links src/PhyloAcc-ST-GBGC/SRC and dumps BPP_C::getRate and BPP_C::getTransition
(GBGC/SRC/bpp_c.cpp:13-30,1405-1427) as hex floats for a grid of (branch, selection coefficient, gBGC bias).
Run here: python tests/golden/make_gbgc_golden.py   ->   tests/golden/gbgc43/{sim.mod,sim.fa,sim.bed,sim.id,params.cfg,probe.txt}"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from phyloacc_b200 import synth  # noqa: E402

PROBE = os.path.join(ROOT, "oracle", "_ref", "gbgc_probe")
OUT = os.path.join(ROOT, "tests", "golden", "gbgc43")


def main():
    tree = synth.make_tree()
    batch = synth.simulate(tree, C=2, length=60, seed=3, gap_frac=0.0, missing_frac=0.0)
    work = tempfile.mkdtemp(prefix="pa_gbgc_")
    cfg = synth.write_files(batch, work, prefix="sim", params=dict(BURNIN=20, MCMC=40, CHAIN=1, ADAPT_FREQ=10, SEED=2, GAPCHAR="-"))
    script = []
    for node in (0, 3, 17, 42, 43, 60, 83):
        for sel in (-2.0, -0.7, 0.0, 0.35, 1.5):
            for bias in (0.0, 0.4, 1.7, 3.2, 6.5):
                script.append(f"tm {node} {sel!r} {bias!r}")
    sp = os.path.join(work, "script.txt")
    with open(sp, "w") as f:
        f.write("\n".join(script) + "\n")
    r = subprocess.run([PROBE, cfg, sp], capture_output=True, text=True, check=True)
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for fn in ("sim.mod", "sim.fa", "sim.bed", "sim.id"):
        shutil.copy(os.path.join(work, fn), os.path.join(OUT, fn))
    with open(os.path.join(OUT, "probe.txt"), "w") as f:
        f.write(r.stdout)
    shutil.rmtree(work)
    print("gbgc43:", len(script), "matrices,", len(r.stdout), "bytes")


if __name__ == "__main__":
    main()
