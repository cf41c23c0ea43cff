"""PhyloAcc-ST-GBGC groundwork: the restatement of the GC-biased per-branch matrices (oracle/gbgc.py) against dumps of
the UNMODIFIED reference (tests/golden/gbgc43/probe.txt from oracle/_ref/gbgc_probe): getRate bit-for-bit up to the last ulp of exp,
getTransition to 1e-10 relative on log P, and the uniformisation route the CUDA kernel uses against both."""
import os

import numpy as np

from oracle import gbgc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gbgc43", "probe.txt")


def _load():
    with open(GOLD) as f:
        return gbgc.parse_probe(f.read())


def test_get_rate_matches_reference():
    Q0, cases = _load()
    assert len(cases) == 175
    for c in cases:
        got = gbgc.get_rate(Q0, c["sel"], c["bias"])
        assert np.allclose(got, c["rate"], rtol=1e-14, atol=0), (c["sel"], c["bias"])
        # rows of a rate matrix: the reference's diagonal update leaves row sums at zero only up to rounding
        assert np.all(np.abs(got.sum(axis=1)) < 1e-12)


def test_log_transition_matches_reference():
    Q0, cases = _load()
    for c in cases:
        got = gbgc.log_transition(Q0, c["sel"], c["bias"], c["t"])
        want = c["logtm"]
        assert np.all(np.abs(got - want) <= 1e-10 * np.maximum(1.0, np.abs(want))), (c["node"], c["sel"], c["bias"])


def test_uniformised_transition_agrees_with_the_eigen_route():
    Q0, cases = _load()
    for c in cases:
        P = gbgc.transition_uniformised(Q0, c["sel"], c["bias"], c["t"])
        assert np.all(P > 0) and np.allclose(P.sum(axis=1), 1.0, rtol=0, atol=1e-13)
        want = c["logtm"].T  # reference layout is (child, parent)
        got = np.log(P)
        assert np.all(np.abs(got - want) <= 1e-10 * np.maximum(1.0, np.abs(want))), (c["node"], c["sel"], c["bias"])


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cuda_gbgc_transition_matches_reference(native_lib):
    """pa_gbgc_transition (pa_gbgc.cu, uniformisation + squaring, one thread per matrix) against the reference's getRate / getTransition."""
    import phyloacc_b200 as pab

    Q0, cases = _load()
    rate, P = pab.gbgc_transition(Q0, [c["sel"] for c in cases], [c["bias"] for c in cases], [c["t"] for c in cases])
    for k, c in enumerate(cases):
        assert np.allclose(rate[k], c["rate"], rtol=1e-13, atol=0)
        want = c["logtm"].T
        got = np.log(P[k])
        assert np.all(P[k] > 0) and np.all(np.abs(got - want) <= 1e-10 * np.maximum(1.0, np.abs(want))), (c["node"], c["sel"], c["bias"])
    # a long batch: rows stay stochastic, long branches reach the stationary distribution of the biased process
    n = 50000
    rng = np.random.default_rng(1)
    sel, bias, t = rng.uniform(-3, 2, n), rng.uniform(0, 8, n), rng.exponential(0.2, n)
    t[:100] = 500.0
    _, P = pab.gbgc_transition(Q0, sel, bias, t)
    assert np.all(P >= 0) and np.allclose(P.sum(axis=2), 1.0, rtol=0, atol=1e-10)  # up to ~14 squarings for the 500-unit branches
    assert np.allclose(P[:100, 0], P[:100, 3], rtol=1e-9)  # every starting base forgotten after 500 substitutions per site


def _replay_case():
    import json

    from tests._util import Case

    case = Case("gbgc43")
    gdir = os.path.dirname(GOLD)
    with open(os.path.join(gdir, "requests.json")) as f:
        reqs = [r for r in json.load(f) if "integ" not in r]
    with open(GOLD) as f:
        want = gbgc.parse_logliks(f.read())
    assert len(reqs) == len(want) == 24
    return case, reqs, want


def _integ_case():
    import json

    from tests._util import Case

    case = Case("gbgc43")
    with open(os.path.join(os.path.dirname(GOLD), "requests.json")) as f:
        reqs = [r for r in json.load(f) if "integ" in r]
    with open(GOLD) as f:
        dumps = gbgc.parse_integ(f.read())
    assert len(reqs) == len(dumps) == 12
    return case, reqs, dumps


def _class_log_tables(Q0, el, r):
    """log P^T ((child, parent) layout) of the six (Z, B) classes of every branch: k = Z + 3 B."""
    out = {}
    for s in el.nodes[:-1]:
        t = float(el.distances2[s])
        out[s] = [np.log(gbgc.transition_uniformised(Q0, (0.0, r["sel_c"], r["sel_n"])[k % 3], r["gbc"] if k >= 3 else 0.0, t)).T for k in range(6)]
    return out


def test_expected_pairs_and_zb_integration_match_reference():
    """The random-number-free half of a GBGC iteration at given states, under each model's own fixZ: expected (child, parent) base pairs per
    branch (Update_Tg(.,.,.,true)), the six (Z, B) emissions, the 6-state upward messages, log_prob_back[root][0] and the ELBO term."""
    case, reqs, dumps = _integ_case()
    Q0, _ = _load()
    pi = np.asarray(case.tree.pi)
    for r, d in zip(reqs, dumps):
        assert d["resZ"] == r["integ"]
        el = case.element(r["c"])
        # the model's constraints as BPP_C::initMCMC sets them (GBGC/SRC/bpp_c.cpp:143-180) and the 3 x 3 tables of Z
        targets = [case.tree.names.index(x) for x in case.P.TARGETSPECIES.split(";") if x]
        fz = gbgc.gbgc_fixz(el, case.tree.N, case.tree.S, d["resZ"], targets)
        assert all(int(fz[s]) == int(d["fixZ"][s]) for s in el.nodes), (r["c"], d["resZ"])
        mine = gbgc._tm_int(el, case.tree.N, case.model.upper, 0.8, 0.1)  # PREP_GRATE / PREP_LRATE defaults of the probe's BPP
        for s in el.nodes[:-1]:
            w, g = np.asarray(d["tmint"][s]), np.asarray(mine[s])
            fin = np.isfinite(w)
            assert w.shape == g.shape and np.array_equal(fin, np.isfinite(g)) and np.allclose(g[fin], w[fin], rtol=1e-14, atol=0), (r["c"], s)
        post = gbgc.expected_pairs(Q0, pi, el, d["Z"], r["B"], r["sel_c"], r["sel_n"], r["gbc"])
        for s, want in d["post"].items():
            assert np.allclose(post[s], want, rtol=1e-9, atol=1e-12), (r["c"], d["resZ"], s)
            assert abs(post[s].sum() - el.GG) < 1e-8  # every column contributes one (child, parent) pair per branch
        tabs = _class_log_tables(Q0, el, r)
        emis, rootback = gbgc.integrate_zb(post, tabs, el, d["fixZ"], d["tmint"], d["consts"])
        for s, want in d["emis"].items():
            fin = np.isfinite(want)
            assert np.array_equal(fin, np.isfinite(emis[s])) and np.allclose(emis[s][fin], want[fin], rtol=1e-10, atol=1e-9), (r["c"], d["resZ"], s)
        fin = np.isfinite(d["rootback"])
        assert np.allclose(rootback[fin], d["rootback"][fin], rtol=1e-10, atol=1e-8)
        assert abs(rootback[0] - d["loglik"]) <= 1e-10 * abs(d["loglik"])
        # ELBO term: the emission of every branch's CURRENT class, without the prior constants (bpp_c2.cpp:386)
        elbo = 0.0
        for s in el.nodes[:-1]:
            if el.missing[s]:
                continue
            k = d["Z"][s] + 3 * r["B"][s]
            if d["fixZ"][s] == 1 and k % 3 == 2:
                continue
            elbo += float(np.where(post[s] == 0, 0.0, tabs[s][k] * post[s]).sum())
        assert abs(elbo - d["elbo"]) <= 1e-10 * abs(d["elbo"])
    case.close()


def test_replayed_gbgc_likelihoods_match_reference():
    """oracle/gbgc.py: replay_loglik against the reference's init_cache_TM + Update_Tg on 24 (element, Z, B, rates) states."""
    case, reqs, want = _replay_case()
    Q0, _ = _load()
    for k, (r, w) in enumerate(zip(reqs, want)):
        if k % 3:  # every third state through the eigen route (slow in numpy), all through the uniformised one below
            continue
        got = gbgc.replay_loglik(Q0, np.asarray(case.tree.pi), case.element(r["c"]), r["Z"], r["B"], r["sel_c"], r["sel_n"], r["gbc"])
        assert abs(got - w) <= 1e-10 * abs(w), (k, got, w)
    for k, (r, w) in enumerate(zip(reqs, want)):
        got = gbgc.replay_loglik(Q0, np.asarray(case.tree.pi), case.element(r["c"]), r["Z"], r["B"], r["sel_c"], r["sel_n"], r["gbc"], uniformised=True)
        assert abs(got - w) <= 1e-10 * abs(w), (k, got, w)
    case.close()


@pytest.mark.gpu
def test_cuda_gbgc_replay_matches_reference(native_lib):
    """pa_engine_gbgc_replay (replay_kernel<GBGC>: per-branch exp(t Q o f(s + b B)) by uniformisation, fp64 pruning over compressed site
    patterns) against the reference's replayed log-likelihoods: 1e-10 relative (north_star's tolerance for replayed likelihoods)."""
    import phyloacc_b200 as pab

    case, reqs, want = _replay_case()
    eng = pab.Engine.from_config(case.cfg, device=0)
    got = eng.gbgc_replay([r["c"] for r in reqs], [r["Z"] for r in reqs], [r["B"] for r in reqs], [r["sel_n"] for r in reqs],
                          [r["sel_c"] for r in reqs], [r["gbc"] for r in reqs])
    for k, (g, w) in enumerate(zip(got, want)):
        assert abs(g - w) <= 1e-10 * abs(w), (k, g, w)
    # B = 0 everywhere and zero selection is the ST model at rates 1: the GBGC replay must agree with the ST replay kernel
    Z = np.zeros((1, eng.N), dtype=np.int8)
    a = eng.gbgc_replay([0], Z, Z, [0.0], [0.0], [0.0])[0]
    b = eng.loglik_replay([0], Z, [1.0], [1.0])[0]
    assert abs(a - b) <= 1e-10 * abs(b)
    eng.close()
    case.close()


@pytest.mark.gpu
def test_cuda_gbgc_integrate_matches_reference(native_lib):
    """pa_engine_gbgc_integrate (gbgc_integrate_kernel: upward pass, downward pass of expected base pairs over compressed site patterns,
    six-class emissions, 6-state messages) against the reference's dumps at twelve states under the three models' own fixZ."""
    import phyloacc_b200 as pab

    case, reqs, dumps = _integ_case()
    eng = pab.Engine.from_config(case.cfg, device=0)
    post, emis, rootv = eng.gbgc_integrate([r["c"] for r in reqs], [d["Z"] for d in dumps], [r["B"] for r in reqs], [d["fixZ"] for d in dumps],
                                           [r["sel_n"] for r in reqs], [r["sel_c"] for r in reqs], [r["gbc"] for r in reqs],
                                           grate=0.8, lrate=0.1, prior_gb=dumps[0]["consts"][0], cons_to_mis=dumps[0]["consts"][1],
                                           ncons_to_mis=dumps[0]["consts"][2])
    for k, (r, d) in enumerate(zip(reqs, dumps)):
        for s, want in d["post"].items():
            assert np.allclose(post[k, s], want, rtol=1e-9, atol=1e-12), (k, s)
        for s, want in d["emis"].items():
            fin = np.isfinite(want)
            assert np.array_equal(fin, np.isfinite(emis[k, s])) and np.allclose(emis[k, s][fin], want[fin], rtol=1e-10, atol=1e-9), (k, s)
        fin = np.isfinite(d["rootback"])
        assert np.allclose(rootv[k, :6][fin], d["rootback"][fin], rtol=1e-10, atol=1e-8), k
        assert abs(rootv[k, 0] - d["loglik"]) <= 1e-10 * abs(d["loglik"])
        assert abs(rootv[k, 6] - d["elbo"]) <= 1e-10 * abs(d["elbo"])
    eng.close()
    case.close()
