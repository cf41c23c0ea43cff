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
