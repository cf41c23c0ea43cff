#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference code.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
It needs oracle/_ref/ref_probe and oracle/_ref/PhyloAcc-ST-ref (built by `make -C oracle ref` from the reference
sources where they lie).  Outputs, all small and committed:

  tests/golden/<case>/{sim.mod,sim.fa,sim.bed,sim.id,params.json}   inputs (synthetic, seeded; phyloacc_b200.synth)
  tests/golden/<case>/probe.json     reference internals via ref_probe: tree arrays, per-element front-end,
                                     replayed log-likelihoods (hex floats), emissions / Z messages for given
                                     histories, Eval2 results for given traces
  tests/golden/<case>/refrun_seed*.json   outputs of the reference binary (elem_lik, rate_postZ) for several seeds

The GPU box has no /root/reference: tests only read these files.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import coracle, frontend as fe, probe  # noqa: E402
from phyloacc_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # 43-taxon ratite-shaped tree, gaps, N, ambiguity codes and missing species (pruned / re-rooted trees occur)
    "ratite43": dict(S=43, tree_seed=43, C=10, length=120, seed=7, gap=0.02, missing=0.08, ambig=0.004,
                     params=dict(BURNIN=20, MCMC=60, CHAIN=1, ADAPT_FREQ=25, SEED=3), n_states=2, chains=[0, 3, 5], refrun_seeds=[1, 2, 3]),
    # small tree, trimming + PRUNE_TREE + Dollo on
    "tiny8": dict(S=8, tree_seed=5, C=6, length=80, seed=11, gap=0.05, missing=0.15, ambig=0.01,
                  params=dict(BURNIN=10, MCMC=40, CHAIN=1, ADAPT_FREQ=10, SEED=2, TRIM_GAP_PERCENT=0.6, PRUNE_TREE=1, HYPER_LRATE2_A=0,
                              MIN_LEN=20, GAP_PROP=0.5),
                  n_states=2, chains=[0, 1], refrun_seeds=[1, 2]),
    # longer chains, several seeds of the reference binary: distributional targets (BF1/BF2, posterior Z, rate medians)
    "post43": dict(S=43, tree_seed=43, C=6, length=100, seed=21, gap=0.01, missing=0.03, ambig=0.0,
                   params=dict(BURNIN=200, MCMC=400, CHAIN=1, ADAPT_FREQ=100, SEED=1), n_states=1, chains=[],
                   refrun_seeds=[1, 2, 3, 4, 5, 6], only=None),
    # BASELINE cfg1 shape (simu_500_200: 43 taxa, 200 bp, 7 outgroup / 9 target / 27 conserved tips) at the DRIVER's default chain
    # lengths (BURNIN 500, MCMC 1000), 100 elements, six seeds of the reference binary with VERBOSE traces: the sampler-behaviour
    # targets of tests/test_gpu_cfg1.py (per-element z-scores, call set, KS on thinned c / n traces).  No probe dump.
    "cfg1_100": dict(S=43, tree_seed=43, n_out=7, n_target=9, C=100, length=200, seed=1, gap=0.01, missing=0.03, ambig=0.0,
                     params=dict(BURNIN=500, MCMC=1000, CHAIN=1, ADAPT_FREQ=500, SEED=1), n_states=0, chains=[], probe=False,
                     refrun_seeds=[1, 2, 3, 4, 5, 6], trace_elements=list(range(12)), trace_thin=10, gz=True, threads=7),
}


def write_case_inputs(name, spec, outdir):
    tree = synth.make_tree(S=spec["S"], seed=spec["tree_seed"], n_out=spec.get("n_out", max(2, spec["S"] // 6)),
                           n_target=spec.get("n_target", max(2, spec["S"] // 5)))
    batch = synth.simulate(tree, C=spec["C"], length=spec["length"], seed=spec["seed"], gap_frac=spec["gap"],
                           missing_frac=spec["missing"], ambig_frac=spec["ambig"])
    cfg = synth.write_files(batch, outdir, prefix="sim", params=spec["params"])
    return tree, batch, cfg


def load_case(outdir):
    """What every test does to get (model, X, starts, lens) from a case directory through the oracle's parsers."""
    P = fe.load_params(os.path.join(outdir, "sim.cfg"))
    tree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    bn, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(tree, P)
    X = fe.reorder_alignment(tree, names, X)
    return P, tree, model, X, starts, lens


def main():
    rng = np.random.default_rng(2026)
    only = sys.argv[1:] or None
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        cdir = os.path.join(GOLD, name)
        if os.path.exists(cdir):
            shutil.rmtree(cdir)
        os.makedirs(cdir)
        work = tempfile.mkdtemp(prefix="pa_gold_")
        tree_s, batch, cfg = write_case_inputs(name, spec, work)
        for f in ("sim.mod", "sim.fa", "sim.bed", "sim.id"):
            shutil.copy(os.path.join(work, f), os.path.join(cdir, f))
        # params without absolute paths; tests re-create the cfg in a temp dir
        with open(cfg) as f:
            lines = [l for l in f.read().split("\n") if l and not l.split()[0].endswith("_FILE") and l.split()[0] != "RESULT_FOLDER"]
        with open(os.path.join(cdir, "params.cfg"), "w") as f:
            f.write("\n".join(lines) + "\n")

        P, tree, model, X, starts, lens = load_case(work)
        N = tree.N
        script = []
        requests = []
        for c in range(spec["C"] if spec.get("probe", True) else 0):
            el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
            script.append(f"element {c}")
            if el.filter:
                continue
            for resZ in (0, 2, 1):
                fixZ = fe.fixz_for_model(el, N, resZ)
                for k in range(spec["n_states"]):
                    Z = -np.ones(N, dtype=np.int8)
                    for nd in reversed(el.nodes):  # top-down so that the Z chain is admissible (no 0 below 1/2, no 2 under 0)
                        if nd == el.root:
                            Z[nd] = rng.integers(0, 2)
                            continue
                        zp = Z[el.parent2[nd]]
                        if zp == 0:
                            opts = [0, 1]
                        elif zp == 1:
                            opts = [1] if fixZ[nd] == 1 else [1, 2]
                        else:
                            opts = [2] if P.HYPER_LRATE2_A == 0 else [1, 2]
                            if fixZ[nd] == 1:
                                opts = [1]
                        Z[nd] = opts[rng.integers(0, len(opts))]
                    nr = float(np.round(rng.gamma(10, 0.2), 6))
                    cr = float(np.round(rng.gamma(5, 0.04), 6))
                    script.append(f"state {resZ} {nr!r} {cr!r} " + " ".join(map(str, Z)))
                    req = dict(c=c, resZ=resZ, Z=Z.tolist(), n_rate=nr, c_rate=cr)
                    if k == 0:
                        # a random ancestral history for the emission / Z-message check
                        g, l, l2 = 0.4, 0.25, (0.0 if P.HYPER_LRATE2_A == 0 else 0.15)
                        hist = rng.integers(0, 4, size=(el.GG, N - tree.S)).astype(np.int8)
                        script.append(f"history {g!r} {l!r} {l2!r}")
                        for row in hist:
                            script.append("h " + " ".join(map(str, row)))
                        req["history"] = dict(g=g, l=l, l2=l2, hist=hist.tolist())
                    requests.append(req)
            if c in spec["chains"]:
                # traces from the oracle chain (Philox schedule) -> the reference's own Eval2 on exactly these traces
                h = coracle.ElementHandle(model, el)
                for resZ in (0, 2, 1):
                    tr = h.chain(resZ, P.SEED, c * 4 + resZ)
                    T = P.BURNIN + P.MCMC
                    script.append(f"eval2 {resZ} {T}")
                    for m in range(T):
                        script.append(f"t {float(tr['loglik'][m])!r} {float(tr['n_rate'][m])!r} {float(tr['c_rate'][m])!r} " + " ".join(map(str, tr["Z"][m])))
                    requests.append(dict(c=c, eval2=resZ, trace={k: np.asarray(tr[k]).tolist() for k in ("loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate", "Z")},
                                         oracle_seed=P.SEED, oracle_key=c * 4 + resZ))
        if spec.get("probe", True):
            sp = os.path.join(work, "script.txt")
            with open(sp, "w") as f:
                f.write("\n".join(script) + "\n")
            r = subprocess.run([probe.PROBE, cfg, sp], capture_output=True, text=True, check=True)
            with open(os.path.join(cdir, "probe.txt"), "w") as f:
                f.write(r.stdout)
            with open(os.path.join(cdir, "requests.json"), "w") as f:
                json.dump(requests, f)
            print(name, "probe done:", len(r.stdout), "bytes")

        # full runs of the reference binary for a few seeds (distributional targets)
        for seed in spec["refrun_seeds"]:
            rw = tempfile.mkdtemp(prefix="pa_goldrun_")
            prm = dict(spec["params"])
            prm["SEED"] = seed
            prm["NUM_THREAD"] = spec.get("threads", 8)
            if spec.get("trace_elements"):
                prm["VERBOSE"] = 1  # the reference then writes <prefix>_mcmc_trace_M<m>_<c>.txt for every element (main.cpp:473-493)
            cfg2 = synth.write_files(batch, rw, prefix="sim", params=prm)
            with open(os.path.join(rw, "log.txt"), "w") as lf:
                subprocess.run([probe.REF_BIN, cfg2], stdout=lf, stderr=subprocess.STDOUT, check=True)
            out = {}
            for fn in ("sim_elem_lik.txt", "sim_rate_postZ_M0.txt", "sim_rate_postZ_M1.txt", "sim_rate_postZ_M2.txt",
                       "sim_M0_elem_Z.txt", "sim_M1_elem_Z.txt", "sim_M2_elem_Z.txt"):
                with open(os.path.join(rw, "out", fn)) as f:
                    out[fn] = f.read()
            if spec.get("trace_elements"):
                # post-burn-in n / c traces of a few elements, thinned (autocorrelation), for the KS check
                thin, burn = spec["trace_thin"], spec["params"]["BURNIN"]
                tr = {}
                for c in spec["trace_elements"]:
                    for m in range(3):
                        fn = os.path.join(rw, "out", f"sim_mcmc_trace_M{m}_{c}.txt")
                        if not os.path.exists(fn):
                            continue
                        with open(fn) as f:
                            rows = [l.split("\t") for l in f.read().split("\n")[1:] if l]
                        tr[f"{c}_{m}"] = dict(n=[float(r[1]) for r in rows[burn::thin]], c=[float(r[2]) for r in rows[burn::thin]],
                                              loglik=[float(r[0]) for r in rows[burn::thin]])
                out["traces"] = tr
            if spec.get("gz"):
                import gzip

                with gzip.open(os.path.join(cdir, f"refrun_seed{seed}.json.gz"), "wt") as f:
                    json.dump(out, f)
            else:
                with open(os.path.join(cdir, f"refrun_seed{seed}.json"), "w") as f:
                    json.dump(out, f)
            shutil.rmtree(rw)
            print(name, "reference run seed", seed, "done")
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
