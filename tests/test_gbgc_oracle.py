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
        reqs = json.load(f)
    with open(GOLD) as f:
        want = gbgc.parse_logliks(f.read())
    assert len(reqs) == len(want) == 24
    return case, reqs, want


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
