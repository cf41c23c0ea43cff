"""Per-chain summaries on the device (pa_summary.cu: MonitorChain's arg-max, Output_init's medians and Z frequencies, Eval2) against
the host implementation of the same estimators (pa_eval.cpp, pinned against the reference's own Eval2 through the oracle, see
test_oracle_pinning.py / test_gpu_chain.py) on the SAME chains: the run is repeated with PA_HOST_SUMMARIES=1, which downloads
every trace and summarises it on the host as round 1 did.  Integer results must be identical, order statistics identical,
Eval2 within 1e-9 (north_star's tolerance for it).  Also: what crosses PCIe."""
import os

import numpy as np
import pytest

import phyloacc_b200 as pab
from phyloacc_b200 import synth
from tests._util import Case

pytestmark = pytest.mark.gpu


def _run(make_engine, ids, host):
    if host:
        os.environ["PA_HOST_SUMMARIES"] = "1"
    else:
        os.environ.pop("PA_HOST_SUMMARIES", None)
    try:
        eng = make_engine()
        res = eng.run(ids)
        cnt = eng.counters()
        out = dict(filtered=res.filtered.copy(), loglik=res.loglik.copy(), loglik_max=res.loglik_max.copy(), max_Z=res.max_Z.copy(),
                   med=res.median_rates.copy(), post_Z=res.post_Z.copy(), cnt=cnt)
        eng.close()
        return out
    finally:
        os.environ.pop("PA_HOST_SUMMARIES", None)


def _compare(dev, host):
    assert np.array_equal(dev["filtered"], host["filtered"])
    assert np.array_equal(dev["max_Z"], host["max_Z"])
    assert np.array_equal(dev["med"], host["med"])          # exact order statistics of the same traces
    assert np.array_equal(dev["post_Z"], host["post_Z"])    # counts / num_mcmc
    a, b = dev["loglik_max"], host["loglik_max"]
    assert np.all((a == b) | (np.abs(a - b) <= 1e-12 * np.abs(b)))
    a, b = dev["loglik"], host["loglik"]
    same = (a == b) | (np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= 1e-9 * np.maximum(1.0, np.abs(b)))
    assert same.all(), (a[~same], b[~same])


@pytest.mark.parametrize("name,extra", [("ratite43", {}), ("tiny8", {}), ("ratite43", dict(BURNIN=100, MCMC=300, ADAPT_FREQ=50)),
                                         ("post43", {})])
def test_device_summaries_equal_host_summaries_on_golden_cases(native_lib, name, extra):
    case = Case(name, extra_params=extra or None)
    ids = list(range(case.C))
    dev = _run(lambda: pab.Engine.from_config(case.cfg, device=0), ids, host=False)
    host = _run(lambda: pab.Engine.from_config(case.cfg, device=0), ids, host=True)
    _compare(dev, host)
    case.close()


def test_device_summaries_on_ragged_batch_and_pcie_bytes(native_lib):
    """cfg2-shaped ragged lengths with gaps, missing species and ambiguity codes; short chains make Eval2 take its pruning fallback on
    some chains (log-det of a degenerate sample covariance), which runs on the host for exactly those."""
    tree = synth.make_tree()
    batch = synth.simulate(tree, C=160, length="cfg2", seed=31, gap_frac=0.02, missing_frac=0.05, ambig_frac=0.002)
    prm = pab.default_params()
    prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed, prm.num_thread = 60, 140, 50, 9, 8

    def make():
        e = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                              conserve=tree.conserve, device=0)
        e.load_alignment(batch.X, batch.elem_start, batch.elem_len)
        return e

    ids = np.arange(batch.C, dtype=np.int32)
    dev = _run(make, ids, host=False)
    host = _run(make, ids, host=True)
    _compare(dev, host)
    n_chains = int((dev["filtered"] == 0).sum()) * 3
    T, N = 200, tree.N
    trace_bytes = n_chains * T * (6 * 8 + ((N + 15) // 16) * 4)
    assert host["cnt"]["d2h_bytes"] >= trace_bytes
    # device path: one record per chain (+ the traces of the chains that took Eval2's host fallback)
    per_chain = 112 + 2 * 4 * ((N + 15) // 16) + 6 * N + 16
    assert dev["cnt"]["d2h_bytes"] <= n_chains * per_chain + dev["cnt"]["host_eval_chains"] * T * (6 * 8 + ((N + 15) // 16) * 4) + 4096
    assert dev["cnt"]["host_eval_chains"] < n_chains


def test_full_length_chains_cross_pcie_as_records_only(native_lib):
    """Driver-default chain lengths: 1500 rows per chain stay in HBM; the host receives < 1 KB per chain."""
    tree = synth.make_tree()
    batch = synth.simulate(tree, C=96, length=200, seed=5)
    prm = pab.default_params()
    prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed, prm.num_thread = 500, 1000, 500, 3, 8
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup, conserve=tree.conserve, device=0)
    eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    res = eng.run(np.arange(batch.C, dtype=np.int32))
    cnt = eng.counters()
    n_chains = int((res.filtered == 0).sum()) * 3
    assert np.isfinite(res.loglik[res.filtered == 0]).all()
    assert cnt["d2h_bytes"] < 1024 * n_chains + cnt["host_eval_chains"] * 1500 * 72
    # a trace asked for afterwards is still there (downloaded on demand)
    tr = eng.trace(0, pab.M2)
    assert tr["loglik"].shape == (1500,) and np.isfinite(tr["loglik"]).all()
    assert eng.counters()["d2h_bytes"] > cnt["d2h_bytes"]
    eng.close()
