#!/usr/bin/env python
"""Generate tests/golden/<case>/hyper.json: the reference's OWN BPP::sample_hyperparam (bpp.cpp:324-477), run through
oracle/_ref/ref_probe ("hyper" command) on seeded per-element rates.  Run here (the container that has /root/reference):
    python tests/golden/make_hyper_golden.py
Each sweep records the hyper-parameters before / after, the rates, and the reference's printed MH ratios and proposals
(17 significant digits).  The random numbers are the compat shim's, so only the arithmetic is pinned: the tests rebuild
the ratios from the recorded proposals."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import probe  # noqa: E402
from tests._util import GOLD, Case  # noqa: E402

FIELDS = ("nprior_a", "nprior_b", "cprior_a", "cprior_b", "prior_l_a", "prior_l_b", "prior_g_a", "prior_g_b", "prior_l2_a", "prior_l2_b")


def main():
    for name, seed in (("ratite43", 11), ("tiny8", 12)):
        case = Case(name)
        rng = np.random.default_rng(seed)
        script, sweeps = [], []
        for it in range(3):
            rates = dict(n=rng.gamma(10, 0.2, case.C), c=rng.gamma(5, 0.04, case.C), g=rng.beta(3, 1, case.C), l=rng.beta(1, 9, case.C),
                         l2=(np.zeros(case.C) if case.P.HYPER_LRATE2_A == 0 else rng.beta(1, 9, case.C)))
            script.append(f"hyper {it} {case.C}")
            for c in range(case.C):
                script.append("r " + " ".join(repr(float(rates[k][c])) for k in ("n", "c", "g", "l", "l2")))
            sweeps.append(dict(iter=it, rates={k: [float(x) for x in v] for k, v in rates.items()}))
        sp = os.path.join(case.dir, "hyper_script.txt")
        with open(sp, "w") as f:
            f.write("\n".join(script) + "\n")
        r = subprocess.run([probe.PROBE, case.cfg, sp], capture_output=True, text=True, check=True)
        k = -1
        for line in r.stdout.split("\n"):
            t = line.split()
            if not t:
                continue
            if t[0] == "hyper_before":
                k += 1
                sweeps[k]["before"] = dict(zip(FIELDS, map(float, t[1:])))
                sweeps[k]["log"] = {}
            elif t[0] == "hyper_after":
                sweeps[k]["after"] = dict(zip(FIELDS, map(float, t[1:])))
            elif t[0] == "hyper_log":
                key = t[1].rstrip(":")
                vals = [float(x.rstrip(",")) for x in t[2:]]
                sweeps[k]["log"][key] = vals
        assert k == len(sweeps) - 1 and all("after" in s for s in sweeps)
        with open(os.path.join(GOLD, name, "hyper.json"), "w") as f:
            json.dump(sweeps, f)
        print(name, "hyper sweeps:", len(sweeps), {kk: len(v) for kk, v in sweeps[0]["log"].items()})
        case.close()


if __name__ == "__main__":
    main()
