"""The GBGC variant's Eval2 (host code of pa_engine_gbgc_run, exposed as pa_gbgc_eval2) against the REFERENCE'S OWN Eval2
(src/PhyloAcc-ST-GBGC/SRC/bpp_c2.cpp:29-285, driven through oracle/_ref/gbgc_probe `eval2`) on the same chains: numpy-oracle chains of
two elements under the three models, plus edited traces that force the estimator's other branches (a single Z configuration; no
configuration above the 5 % weight: the q(r) q(Z) fallback; configurations seen fewer than 11 times: diagonal covariance).  The same chains
go through the reference's own Output_init (the row of _rate_postZ_<resZ>.txt: "medians" of the three rates, P(B), P(Z) per node with its
rule for missing nodes) and through pa_gbgc_output_init: equal at the file's 6 significant digits.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from oracle import gbgc
from tests._util import Case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "oracle", "_ref", "gbgc_probe")
GOLD = os.path.join(ROOT, "tests", "golden")
MODELS = (0, 2, 1)


def _reference_eval2(case, c, requests, tmp):
    """requests: [(resZ, burn, trace dict)] -> ([Eval2 estimate], [the row Output_init writes, split at tabs])"""
    lines = [f"element {c}"]
    for k, (resZ, burn, tr) in enumerate(requests):
        path = os.path.join(tmp, f"trace_{c}_{k}.txt")
        T = len(tr["full"])
        with open(path, "w") as f:
            for m in range(T):
                f.write(" ".join([repr(float(tr[k][m])) for k in ("full", "c", "n", "gb")] +  # repr round-trips a double exactly (operator>> reads no hex floats)
                                 [str(int(z)) for z in tr["Z"][m]] + [str(int(b)) for b in tr["B"][m]]) + "\n")
        lines.append(f"eval2 {resZ} {burn} {T - burn} {path}")
        lines.append(f"output_init {resZ} {burn} {T - burn} {path}")
    script = os.path.join(tmp, f"script_{c}.txt")
    with open(script, "w") as f:
        f.write("\n".join(lines) + "\n")
    r = subprocess.run([PROBE, case.cfg, script], capture_output=True, text=True, check=True)
    rows = [l.split("\t")[1:] for l in r.stdout.split("\n") if l.startswith("output_init ")]
    return [float.fromhex(l.split()[2]) for l in r.stdout.split("\n") if l.startswith("eval2 ")], rows


@pytest.mark.skipif(not os.path.exists(PROBE), reason="oracle/_ref/gbgc_probe not built (make -C oracle ref-gbgc)")
def test_host_eval2_matches_the_reference(native_lib, tmp_path):
    import phyloacc_b200 as pab

    case = Case("gbgc_chain43")
    with open(os.path.join(GOLD, "gbgc43", "probe.txt")) as f:
        Q0, _ = gbgc.parse_probe(f.read())
    tree = case.tree
    targets = [tree.names.index(x) for x in case.P.TARGETSPECIES.split(";") if x]
    Q0 = np.asarray(tree.Q, float).reshape(4, 4)
    burn, mcmc = 30, 120
    n_checked = 0
    for c in (1, 9):
        el = case.element(c)
        reqs = []
        for mdl in range(3):
            par = gbgc.GbgcParams(num_burn=burn, num_mcmc=mcmc)
            tr = gbgc.GbgcChain(Q0, tree.pi, el, tree.N, tree.S, MODELS[mdl], targets, case.model.upper, par, 41 + c, c * 4 + MODELS[mdl]).run()
            tr = {k: np.array(tr[k]) for k in ("full", "c", "n", "gb", "Z", "B")}
            tr["Z"] = np.where(tr["Z"] < 0, 1, tr["Z"])  # pruned nodes: the reference's trace_Z keeps its initial 1 there
            reqs.append((MODELS[mdl], burn, tr))
            if mdl == 2:
                one = {k: v.copy() for k, v in tr.items()}  # a single Z configuration: every sample equals the arg-max state
                one["Z"][:] = one["Z"][burn + int(np.argmax(one["full"][burn:]))]
                reqs.append((1, burn, one))
                many = {k: v.copy() for k, v in tr.items()}  # no configuration reaches 5 %: every sample its own configuration
                node = el.nodes[0]
                rng = np.random.default_rng(c)
                for m in range(len(many["full"])):
                    many["Z"][m, el.nodes[:7]] = [(m >> b) & 1 for b in range(7)]
                reqs.append((1, burn, many))
                few = {k: v.copy() for k, v in tr.items()}  # configurations of exactly 8 samples each: diagonal covariance (<= 10), weight 8/120 > 5 %
                for m in range(len(few["full"])):
                    few["Z"][m, el.nodes[:4]] = [((m // 8) >> b) & 1 for b in range(4)]
                reqs.append((1, burn, few))
        want, rows = _reference_eval2(case, c, reqs, str(tmp_path))
        assert len(want) == len(reqs) == len(rows)
        for (resZ, b0, tr), row in zip(reqs, rows):
            # the reference's own _rate_postZ row for this chain (Output_init) against pa_gbgc_output_init, at the file's 6 significant digits
            med, pb, pz = pab.gbgc_output_init(tr["n"], tr["c"], tr["gb"], tr["Z"], tr["B"], el.missing, b0)
            mine = [f"{x:.6g}" for x in np.concatenate([med, pb, pz.ravel()])]
            assert row[0] == str(c) and [f"{float(x):.6g}" for x in row[1:]] == mine, (c, resZ)
        for (resZ, b0, tr), w in zip(reqs, want):
            got, _ = pab.gbgc_eval2(tr["full"], tr["c"], tr["n"], tr["gb"], tr["Z"], tr["B"], el.nodes, el.missing, b0, resZ)
            assert np.isfinite(w) == np.isfinite(got)
            if np.isfinite(w):
                assert abs(got - w) <= 1e-9 * max(1.0, abs(w)), (c, resZ, got, w)
            n_checked += 1
    assert n_checked == 12
    case.close()


@pytest.mark.skipif(not os.path.exists(PROBE), reason="oracle/_ref/gbgc_probe not built (make -C oracle ref-gbgc)")
def test_recorded_full_loglik_and_prior_of_z_match_the_reference(tmp_path):
    """What a chain RECORDS for a state: log P(constraints) of initMCMC, logp_Z = prior_Z(Z) - log P(constraints) and MonitorChain's
    log P(X, Z, B, r).  The CUDA chain is compared draw for draw with oracle/gbgc.py; here the numpy formulas are pinned against the
    reference's own initMCMC / prior_Z / MonitorChain (gbgc_probe `monitor`) at the states a short chain visits, under the three models."""
    case = Case("gbgc_chain43")
    tree = case.tree
    targets = [tree.names.index(x) for x in case.P.TARGETSPECIES.split(";") if x]
    Q0 = np.asarray(tree.Q, float).reshape(4, 4)
    n = 0
    for c in (0, 7):
        el = case.element(c)
        lines, want = [f"element {c}"], []
        for mdl in range(3):
            par = gbgc.GbgcParams(num_burn=4, num_mcmc=14)
            ch = gbgc.GbgcChain(Q0, tree.pi, el, tree.N, tree.S, MODELS[mdl], targets, case.model.upper, par, 77, c * 4 + MODELS[mdl])
            tr = ch.run()
            for m in (0, 5, 11, 17):
                Z, B = tr["Z"][m], tr["B"][m]
                lines.append(f"monitor {MODELS[mdl]} {float(tr['loglik'][m])!r} {float(tr['c'][m])!r} {float(tr['n'][m])!r} {float(tr['gb'][m])!r} " +
                             " ".join(str(int(z)) for z in Z) + " " + " ".join(str(int(b)) for b in B))
                want.append((float(tr["full"][m]), float(ch.prior_z(Z) - tr["logp_zc"]), float(tr["logp_zc"])))
        script = os.path.join(str(tmp_path), f"mon_{c}.txt")
        with open(script, "w") as f:
            f.write("\n".join(lines) + "\n")
        r = subprocess.run([PROBE, case.cfg, script], capture_output=True, text=True, check=True)
        got = [[float.fromhex(x) for x in l.split()[2:5]] for l in r.stdout.split("\n") if l.startswith("monitor ")]
        assert len(got) == len(want) == 12
        for g, w in zip(got, want):
            assert abs(g[2] - w[2]) <= 1e-12 * abs(w[2]) and abs(g[1] - w[1]) <= 1e-12 * max(1.0, abs(w[1])), (g, w)
            assert abs(g[0] - w[0]) <= 1e-12 * abs(w[0]), (g, w)
            n += 1
    assert n == 24
    case.close()
