"""Size-independent properties of the on-device chain at the bench workload's sizes (BASELINE config 2 shape: 43 taxa,
element lengths up to 2000 bp, several pattern tiles per element), where the plain-C oracle would take minutes:

* replay consistency: the log-likelihood the chain carries through its incremental passes (dirty ancestors only, two
  arena copies with per-node selectors, register forwarding / stack / arena operands) equals a FRESH full pruning of the
  recorded state (Z, n, c) by the independent replay kernel, to 1e-10 relative, at every sampled iteration;
* the sampler's hard constraints hold in every trace row: root Z in {0, 1}; no state 2 under M0 or where fixZ == 1;
  a node is 0 only under a 0 parent... (Dollo: with HYPER_LRATE2_A == 0 no return from 2 to 1); rate bounds of ROPT 1;
* summaries are consistent: posterior Z rows sum to one, medians lie in the traces, logBF columns follow the log-likelihoods.
"""
import numpy as np
import pytest

import phyloacc_b200 as pab
from phyloacc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run(native_lib):
    tree = synth.make_tree()
    rng = np.random.default_rng(5)
    lengths = [2000, 1500, 1100, 777, 513, 300, 257, 256, 255, 97] + list(rng.integers(50, 400, size=22))
    batch = synth.simulate(tree, C=len(lengths), length=200, seed=31, lengths=lengths, gap_frac=0.02, missing_frac=0.05, ambig_frac=0.002)
    prm = pab.default_params()
    prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed, prm.num_thread = 20, 40, 10, 3, 8
    prm.hyper_lrate2_a = 0.0  # Dollo (the driver's default, main.cpp:286)
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                            conserve=tree.conserve, device=0)
    eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    ids = np.arange(batch.C, dtype=np.int32)
    res = eng.run(ids)
    yield tree, batch, prm, eng, ids, res
    eng.close()


def test_trace_loglik_equals_fresh_pruning_of_the_recorded_state(run):
    tree, batch, prm, eng, ids, res = run
    T = prm.burnin + prm.mcmc
    rows = [0, 1, 7, 23, T - 1]
    n_checked = 0
    for i, c in enumerate(ids):
        if res.filtered[i]:
            continue
        for mdl in (pab.M0, pab.M1, pab.M2):
            tr = eng.trace(i, mdl)
            Z = np.stack([tr["Z"][m] for m in rows])
            got = eng.loglik_replay([int(c)] * len(rows), Z, [float(tr["n_rate"][m]) for m in rows], [float(tr["c_rate"][m]) for m in rows])
            want = np.array([tr["loglik"][m] for m in rows])
            assert np.all(np.abs(got - want) <= 1e-10 * np.abs(want)), (int(c), mdl, got, want)
            n_checked += len(rows)
    assert n_checked >= 3 * 5 * 20


def test_hard_constraints_hold_in_every_trace_row(run):
    tree, batch, prm, eng, ids, res = run
    N = tree.N
    parent = np.asarray(tree.parent)
    for i, c in enumerate(ids):
        if res.filtered[i]:
            continue
        info = eng.element_info(int(c), gg_max=2048)
        nodes = np.array(info["nodes"][: info["n_nodes"]])
        root = info["root"]
        par2 = np.asarray(info["parent2"])
        for mdl in (pab.M0, pab.M1, pab.M2):
            tr = eng.trace(i, mdl)
            Z = tr["Z"]  # [T][N], -1 outside the pruned tree
            assert (Z[:, nodes] >= 0).all() and (Z[:, nodes] <= 2).all()
            pruned_out = np.setdiff1d(np.arange(N), nodes)
            assert (Z[:, pruned_out] == -1).all()
            assert (Z[:, root] <= 1).all()  # root prior (1/2, 1/2, 0), bpp_c.hpp:197-200
            fix = np.asarray(info["fixZ"][mdl]) == 1
            assert not (Z[:, fix & np.isin(np.arange(N), nodes)] == 2).any()
            if mdl == pab.M0:
                assert not (Z[:, nodes] == 2).any()
            for s in nodes:
                if s == root:
                    continue
                zp, zc = Z[:, par2[s]], Z[:, s]
                assert not ((zc == 0) & (zp != 0)).any()      # state 0 (background) is only entered at the root side
                assert not ((zc == 1) & (zp == 2)).any()      # Dollo: no return from accelerated to conserved
                assert not ((zc == 2) & (zp == 0)).any()      # 0 -> 2 has probability 0 (bpp_c.cpp:301-332)
            assert (tr["c_rate"] <= prm.cub + 1e-15).all() and (tr["c_rate"] > 0).all()
            if mdl != pab.M0:
                assert (tr["n_rate"] >= prm.nlb - 1e-15).all()
            assert np.isfinite(tr["loglik"]).all()


def test_summaries_are_consistent_with_the_traces(run):
    tree, batch, prm, eng, ids, res = run
    T = prm.burnin + prm.mcmc
    for i, c in enumerate(ids):
        if res.filtered[i]:
            assert (res.loglik[i] == 0).all()
            continue
        info = eng.element_info(int(c), gg_max=2048)
        miss = np.asarray(info["missing"]).astype(bool)
        rows = res.post_Z[i].sum(axis=-1)  # [3][N]
        assert np.allclose(rows[:, ~miss], 1.0)
        # reference quirk kept (bpp_c2.cpp:764-772): a missing node's column 0 is forced to 1 on top of its sampled frequencies
        assert np.allclose(res.post_Z[i][:, miss, 0], 1.0)
        for mdl in (pab.M0, pab.M1, pab.M2):
            tr = eng.trace(i, mdl)
            for k, col in enumerate(("n_rate", "c_rate", "g_rate", "l_rate", "l2_rate")):
                assert res.median_rates[i, mdl, k] == np.sort(tr[col])[T // 2]
            assert res.loglik_max[i, mdl] <= 0 or np.isfinite(res.loglik_max[i, mdl])
    ok = np.isfinite(res.loglik).all(axis=1) & (res.filtered == 0)
    assert np.allclose(res.logBF1[ok], res.loglik[ok, pab.M1] - res.loglik[ok, pab.M0])
    assert np.allclose(res.logBF2[ok], res.loglik[ok, pab.M1] - res.loglik[ok, pab.M2])
