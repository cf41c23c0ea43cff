#!/usr/bin/env python
"""Chain-level fixture for PhyloAcc-ST-GBGC: the UNMODIFIED reference binary (oracle/_ref/PhyloAcc-GBGC-ref, built by
`make -C oracle ref-gbgc` from src/PhyloAcc-ST-GBGC/SRC against the Armadillo / GSL stand-ins) run with SIX seeds on 12 elements of
43 taxa x 200 bp at the variant's default chain lengths (BURNIN 200 + MCMC 800), VERBOSE traces on.  Per seed, element and model the
fixture keeps: the medians of n_rate / c_rate / gBC, mean and s.d. of the post-burn-in log P(X, Z, B, r), P(B = 1) and P(Z = k) per node
(from the traces, full precision of the 8-digit print), and the three marginal log-likelihoods of _elem_lik.txt.
Run here (takes ~4 min on 6 cores):  python tests/golden/make_gbgc_chain_golden.py  ->  tests/golden/gbgc_chain43/"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from phyloacc_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-GBGC-ref")
OUT = os.path.join(ROOT, "tests", "golden", "gbgc_chain43")
SEEDS = (11, 12, 13, 14, 15, 16)
BURN, MCMC = 200, 800


def run_seed(work, seed):
    d = os.path.join(work, f"seed{seed}")
    os.makedirs(os.path.join(d, "out"))
    with open(os.path.join(work, "sim.cfg")) as f:
        body = f.read()
    lines = [l for l in body.split("\n") if l and l.split()[0] not in ("SEED", "RESULT_FOLDER")]
    cfg = os.path.join(d, "sim.cfg")
    with open(cfg, "w") as f:
        f.write("\n".join(lines) + f"\nSEED {seed}\nRESULT_FOLDER {d}/out\n")
    subprocess.run([REF, cfg], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    return d


def summarise(d, C, N):
    res = {"lik": {}, "chains": {}}
    rows = open(os.path.join(d, "out", "sim_elem_lik.txt")).read().strip().split("\n")[1:]
    for r in rows:
        f = r.split("\t")
        res["lik"][f[0]] = [float(x) for x in f[2:]]
    for c in range(C):
        for resZ in (0, 2, 1):
            p = os.path.join(d, "out", f"sim_mcmc_trace_{resZ}_{c}.txt")
            if not os.path.exists(p):
                continue
            a = np.array([[float(x) for x in r.split("\t") if x != ""] for r in open(p).read().strip().split("\n")[1:]])
            a = a[BURN:]
            Bm, Zm = a[:, 4:4 + N], a[:, 4 + N:4 + 2 * N]
            res["chains"][f"{c}:{resZ}"] = dict(
                med=[float(np.median(a[:, 1])), float(np.median(a[:, 2])), float(np.median(a[:, 3]))],
                mean=[float(a[:, 1].mean()), float(a[:, 2].mean()), float(a[:, 3].mean())],
                full=[float(a[:, 0].mean()), float(a[:, 0].std())],
                pB=[round(float(x), 5) for x in Bm.mean(axis=0)],
                pZ=[[round(float((Zm[:, s] == k).mean()), 5) for k in range(3)] for s in range(N)],
            )
    return res


def main():
    tree = synth.make_tree()
    batch = synth.simulate(tree, C=12, length=200, seed=5, gap_frac=0.01, missing_frac=0.03, ambig_frac=0.002)
    work = tempfile.mkdtemp(prefix="pa_gbgc_chain_")
    synth.write_files(batch, work, prefix="sim", params=dict(BURNIN=BURN, MCMC=MCMC, CHAIN=1, ADAPT_FREQ=100, SEED=1, GAPCHAR="-", VERBOSE=1, NUM_THREAD=1))
    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        dirs = list(ex.map(lambda s: run_seed(work, s), SEEDS))
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for fn in ("sim.mod", "sim.fa", "sim.bed", "sim.id"):
        shutil.copy(os.path.join(work, fn), os.path.join(OUT, fn))
    with open(os.path.join(work, "sim.cfg")) as f:
        lines = [l for l in f.read().split("\n") if l and not l.split()[0].endswith("_FILE") and l.split()[0] != "RESULT_FOLDER"]
    with open(os.path.join(OUT, "params.cfg"), "w") as f:
        f.write("\n".join(lines) + "\n")
    for seed, d in zip(SEEDS, dirs):
        with gzip.open(os.path.join(OUT, f"refrun_seed{seed}.json.gz"), "wt") as f:
            json.dump(summarise(d, batch.C, tree.N), f)
    shutil.rmtree(work)
    print("gbgc_chain43 written:", os.listdir(OUT))


if __name__ == "__main__":
    main()
