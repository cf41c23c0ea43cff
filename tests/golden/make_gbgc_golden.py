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
    batch = synth.simulate(tree, C=4, length=120, seed=3, gap_frac=0.02, missing_frac=0.0, ambig_frac=0.003)
    work = tempfile.mkdtemp(prefix="pa_gbgc_")
    cfg = synth.write_files(batch, work, prefix="sim", params=dict(BURNIN=20, MCMC=40, CHAIN=1, ADAPT_FREQ=10, SEED=2, GAPCHAR="-"))
    script = []
    for node in (0, 3, 17, 42, 43, 60, 83):
        for sel in (-2.0, -0.7, 0.0, 0.35, 1.5):
            for bias in (0.0, 0.4, 1.7, 3.2, 6.5):
                script.append(f"tm {node} {sel!r} {bias!r}")
    # replayed log-likelihoods: random (Z, B) states and rates for every element (the reference evaluates them through its own
    # init_cache_TM + Update_Tg; tests replay them through pa_engine_gbgc_replay)
    import json

    import numpy as np

    rng = np.random.default_rng(11)
    N = tree.N
    requests = []
    for c in range(batch.C):
        script.append(f"element {c}")
        for k in range(6):
            Z = rng.integers(0, 3, size=N)
            B = rng.integers(0, 2, size=N)
            if k == 0:
                Z[:] = 1
                B[:] = 0
            sc, sn, gb = float(np.round(rng.uniform(-3, -0.5), 6)), float(np.round(rng.uniform(-1, 0.5), 6)), float(np.round(rng.uniform(0, 5), 6))
            script.append(f"state {sc!r} {sn!r} {gb!r} " + " ".join(map(str, Z)) + " " + " ".join(map(str, B)))
            requests.append(dict(c=c, sel_c=sc, sel_n=sn, gbc=gb, Z=Z.tolist(), B=B.tolist()))
        # the random-number-free evaluations of one Gibbs iteration (expected base pairs, emissions, 6-state messages, ELBO) under each
        # model's own fixZ: M0 (resZ 0), M1 (resZ 2), M2 (resZ 1)
        for resZ in (0, 2, 1):
            Z = rng.integers(0, 3, size=N)
            B = rng.integers(0, 2, size=N)
            sc, sn, gb = float(np.round(rng.uniform(-3, -0.5), 6)), float(np.round(rng.uniform(-1, 0.5), 6)), float(np.round(rng.uniform(0, 5), 6))
            script.append(f"integ {resZ} {sc!r} {sn!r} {gb!r} " + " ".join(map(str, Z)) + " " + " ".join(map(str, B)))
            requests.append(dict(c=c, integ=resZ, sel_c=sc, sel_n=sn, gbc=gb, B=B.tolist()))
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
    with open(os.path.join(OUT, "requests.json"), "w") as f:
        json.dump(requests, f)
    with open(cfg) as f:
        lines = [l for l in f.read().split("\n") if l and not l.split()[0].endswith("_FILE") and l.split()[0] != "RESULT_FOLDER"]
    with open(os.path.join(OUT, "params.cfg"), "w") as f:
        f.write("\n".join(lines) + "\n")
    shutil.rmtree(work)
    print("gbgc43:", len(script), "matrices,", len(r.stdout), "bytes")


if __name__ == "__main__":
    main()
