"""PhyloAcc-ST-GBGC chain.  Three layers of evidence, like the ST path:
  * oracle/gbgc.py:GbgcChain (numpy restatement of BPP_C::Gibbs, GBGC/SRC/bpp_c.cpp:259-345) against runs of the UNMODIFIED reference binary
    (tests/golden/gbgc_chain43: six seeds of oracle/_ref/PhyloAcc-GBGC-ref, generator tests/golden/make_gbgc_chain_golden.py) -- statistically,
    the reference's GSL stream is not reproduced;
  * the CUDA chain (pa_engine_gbgc_run, pa_gbgc_chain.cu) against the oracle DRAW FOR DRAW on the shared Philox schedule: identical (Z, B)
    traces, rates and log-likelihoods to 1e-9;
  * the CUDA chain against the reference's seeds at full length: medians of the three rates, P(B) and P(Z) per node, the marginal
    log-likelihoods, as z-scores on the seed-to-seed standard errors of both sides."""
import os

import numpy as np
import pytest

from oracle import gbgc
from tests._util import Case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = (0, 2, 1)  # model index 0 / 1 / 2 = M0 / M1 / M2 -> the reference's resZ


def _setup(case):
    with open(os.path.join(GOLD, "gbgc43", "probe.txt")) as f:
        Q0, _ = gbgc.parse_probe(f.read())  # the same .mod rate matrix as gbgc_chain43 (synth.make_tree defaults)
    tree = case.tree
    assert np.allclose(Q0, np.asarray(tree.Q).reshape(4, 4))
    targets = [tree.names.index(x) for x in case.P.TARGETSPECIES.split(";") if x]
    return np.asarray(tree.Q, float).reshape(4, 4), targets


def _oracle_chain(case, Q0, targets, c, mdl, seed, burn, mcmc):
    par = gbgc.GbgcParams(num_burn=burn, num_mcmc=mcmc)
    ch = gbgc.GbgcChain(Q0, case.tree.pi, case.element(c), case.tree.N, case.tree.S, MODELS[mdl], targets, case.model.upper, par, seed, c * 4 + MODELS[mdl])
    return ch.run()


def _ref_runs(case):
    runs = case.golden_refruns()
    assert len(runs) == 6
    return runs


def test_oracle_chain_agrees_with_the_reference_binary():
    """One element, three models, 100 + 300 iterations of the numpy chain against the six reference seeds (200 + 800): the post-burn-in
    mean of log P(X, Z, B, r), the medians of the conserved selection coefficient and the bias, the expected number of gBGC branches and of
    accelerated branches must lie within the reference's own seed-to-seed range widened by 4 standard errors of the shorter chain."""
    case = Case("gbgc_chain43")
    Q0, targets = _setup(case)
    runs = _ref_runs(case)
    c = 1
    for mdl in range(3):
        tr = _oracle_chain(case, Q0, targets, c, mdl, 17, 100, 300)
        ref = [r["chains"][f"{c}:{MODELS[mdl]}"] for r in runs]
        b = 100
        mine = dict(full=tr["full"][b:].mean(), c=np.median(tr["c"][b:]), gb=np.median(tr["gb"][b:]), nB=tr["B"][b:].sum(axis=1).mean(),
                    nZ2=(tr["Z"][b:] == 2).sum(axis=1).mean())
        theirs = dict(full=[r["full"][0] for r in ref], c=[r["med"][1] for r in ref], gb=[r["med"][2] for r in ref], nB=[sum(r["pB"]) for r in ref],
                      nZ2=[sum(z[2] for z in r["pZ"]) for r in ref])
        for k, v in mine.items():
            t = np.asarray(theirs[k])
            slack = 4 * max(t.std(ddof=1), 1e-3) * np.sqrt(800 / 300) + 0.02 * abs(t.mean())
            assert t.min() - slack <= v <= t.max() + slack, (mdl, k, v, t)
    case.close()


@pytest.mark.gpu
def test_cuda_chain_matches_the_oracle_draw_for_draw(native_lib):
    import phyloacc_b200 as pab

    case = Case("gbgc_chain43")
    Q0, targets = _setup(case)
    eng = pab.Engine.from_config(case.cfg, device=0)
    ids = [0, 3, 7]
    burn, mcmc, seed = 10, 40, 23
    out = eng.gbgc_run(ids, targets, traces=True, num_burn=burn, num_mcmc=mcmc, seed=seed)
    assert out["chain_iterations"] == len(ids) * 3 * (burn + mcmc - 1)
    for i, c in enumerate(ids):
        for mdl in range(3):
            tr = _oracle_chain(case, Q0, targets, c, mdl, seed, burn, mcmc)
            got = out["trace"][i, mdl]
            assert np.array_equal(out["trace_z"][i, mdl], tr["Z"]), (c, mdl)
            assert np.array_equal(out["trace_b"][i, mdl], tr["B"]), (c, mdl)
            for row, key in ((0, "loglik"), (1, "full"), (2, "c"), (3, "n"), (4, "gb")):
                assert np.allclose(got[row], tr[key], rtol=1e-9, atol=1e-9), (c, mdl, key)
            assert out["accept"][i, mdl, 0] == tr["accept_rates"] and out["accept"][i, mdl, 1] == tr["accept_zb"]
    eng.close()
    case.close()


@pytest.mark.gpu
def test_cuda_chain_on_long_elements_matches_the_oracle(native_lib, tmp_path):
    """1500-bp elements: ~1000 site patterns, i.e. eight pattern tiles per pass (the per-warp pair sums accumulate across tiles) and tip codes
    too large for the shared-memory staging buffer (read from the element image instead) -- paths the 200-bp fixtures never take."""
    import phyloacc_b200 as pab
    from oracle import frontend as fe
    from phyloacc_b200 import synth

    stree = synth.make_tree()
    batch = synth.simulate(stree, C=2, length=1500, seed=8, gap_frac=0.01, missing_frac=0.03, ambig_frac=0.002)
    cfg = synth.write_files(batch, str(tmp_path), prefix="sim", params=dict(BURNIN=3, MCMC=6, CHAIN=1, SEED=1, GAPCHAR="-"))
    P = fe.load_params(cfg)
    otree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    _, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(otree, P)
    X = fe.reorder_alignment(otree, names, X)
    eng = pab.Engine.from_config(cfg, device=0)
    burn, mcmc, seed = 3, 6, 31
    out = eng.gbgc_run([0, 1], stree.targets, traces=True, num_burn=burn, num_mcmc=mcmc, seed=seed)
    Q0 = np.asarray(otree.Q, float).reshape(4, 4)
    for c in range(2):
        el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
        assert fe.compress_patterns(el.masks).weights.shape[0] > 600
        for mdl in (0, 2):
            par = gbgc.GbgcParams(num_burn=burn, num_mcmc=mcmc)
            tr = gbgc.GbgcChain(Q0, otree.pi, el, otree.N, otree.S, MODELS[mdl], stree.targets, model.upper, par, seed, c * 4 + MODELS[mdl]).run()
            assert np.array_equal(out["trace_z"][c, mdl], tr["Z"]) and np.array_equal(out["trace_b"][c, mdl], tr["B"]), (c, mdl)
            for row, key in ((0, "loglik"), (1, "full"), (2, "c"), (3, "n"), (4, "gb")):
                assert np.allclose(out["trace"][c, mdl][row], tr[key], rtol=1e-9, atol=1e-9), (c, mdl, key)
    eng.close()


@pytest.mark.gpu
def test_cuda_chain_states_replay_to_their_recorded_likelihoods(native_lib):
    """Every recorded state of a chain, replayed from scratch by pa_engine_gbgc_replay (an independent kernel), gives the recorded
    log P(X | Z, B, r) to 1e-10 (north_star's replay criterion)."""
    import phyloacc_b200 as pab

    case = Case("gbgc_chain43")
    _, targets = _setup(case)
    eng = pab.Engine.from_config(case.cfg, device=0)
    ids = list(range(case.C))
    out = eng.gbgc_run(ids, targets, traces=True, num_burn=20, num_mcmc=60, seed=5)
    rng = np.random.default_rng(3)
    el, Z, B, sn, sc, gb, want = [], [], [], [], [], [], []
    for i in range(len(ids)):
        if out["filtered"][i]:
            continue
        for mdl in range(3):
            for m in rng.choice(80, size=6, replace=False):
                el.append(ids[i]); Z.append(np.maximum(out["trace_z"][i, mdl, m], 0)); B.append(out["trace_b"][i, mdl, m])
                sc.append(out["trace"][i, mdl, 2, m]); sn.append(out["trace"][i, mdl, 3, m]); gb.append(out["trace"][i, mdl, 4, m])
                want.append(out["trace"][i, mdl, 0, m])
    got = eng.gbgc_replay(el, np.asarray(Z), np.asarray(B), sn, sc, gb)
    want = np.asarray(want)
    assert np.all(np.abs(got - want) <= 1e-10 * np.abs(want)), np.abs(got / want - 1).max()
    eng.close()
    case.close()


@pytest.mark.gpu
def test_cuda_chain_matches_the_reference_across_seeds(native_lib):
    """Eight seeds of the CUDA chain against six seeds of the unmodified reference, 12 elements x 3 models at the variant's default chain
    lengths: per (element, model) the medians of the three rates and the mean log P(X, Z, B, r); per node P(B = 1) and P(Z = k); per
    element the three marginal log-likelihoods.  z = difference of the seed means over the pooled standard error (with a floor for
    near-constant quantities); the z distribution must be centred (|mean| < 0.5) and light-tailed (|z| > 5 in under 0.5 % of the entries)."""
    import phyloacc_b200 as pab

    case = Case("gbgc_chain43")
    _, targets = _setup(case)
    runs = _ref_runs(case)
    eng = pab.Engine.from_config(case.cfg, device=0)
    ids = list(range(case.C))
    mine = [eng.gbgc_run(ids, targets, num_burn=200, num_mcmc=800, seed=100 + s) for s in range(8)]
    z_all, lik_z = [], []

    def zscore(a, b, floor):
        a, b = np.asarray(a, float), np.asarray(b, float)
        se = np.sqrt(a.var(axis=0, ddof=1) / a.shape[0] + b.var(axis=0, ddof=1) / b.shape[0])
        return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, floor)

    for i, c in enumerate(ids):
        if mine[0]["filtered"][i]:
            assert f"{c}:0" not in runs[0]["chains"]
            continue
        for mdl in range(3):
            ref = [r["chains"][f"{c}:{MODELS[mdl]}"] for r in runs]
            z_all.append(zscore([m["med"][i, mdl] for m in mine], [r["med"] for r in ref], 0.02))
            z_all.append(zscore([m["post_b"][i, mdl] for m in mine], [r["pB"] for r in ref], 0.01))
            z_all.append(zscore([m["post_z"][i, mdl][:, 1:] for m in mine], [r["pZ"] for r in ref], 0.01).ravel())
        ref_l = np.asarray([r["lik"][str(c)][:3] for r in runs])       # loglik_NUll, loglik_RES, loglik_all
        my_l = np.asarray([m["loglik"][i] for m in mine])
        lik_z.append(zscore(my_l, ref_l, 0.3))
        ref_m = np.asarray([r["lik"][str(c)][4:7] for r in runs])      # loglik_Max of resZ 0, 1, 2
        my_m = np.asarray([m["loglik_max"][i][[0, 2, 1]] for m in mine])
        lik_z.append(zscore(my_m, ref_m, 0.3))
    z = np.concatenate([np.ravel(x) for x in z_all])
    lz = np.concatenate([np.ravel(x) for x in lik_z])
    assert abs(z.mean()) < 0.5 and np.mean(np.abs(z) > 5) < 0.005, (z.mean(), np.mean(np.abs(z) > 5), np.abs(z).max())
    assert abs(lz.mean()) < 1.0 and np.abs(lz).max() < 6, (lz.mean(), np.abs(lz).max())
    eng.close()
    case.close()


@pytest.mark.gpu
def test_cuda_chain_on_a_240_taxon_tree_replays_to_its_recorded_likelihoods(native_lib, tmp_path):
    """Zoonomia-scale tree (479 nodes: one CTA per SM, 212 KB of shared memory): the chain runs and every sampled state, replayed from
    scratch by pa_engine_gbgc_replay (an independent kernel), gives the recorded log P(X | Z, B, r) to 1e-10."""
    import phyloacc_b200 as pab
    from phyloacc_b200 import synth

    S = 240
    stree = synth.make_tree(S=S, seed=9, n_out=S // 6, n_target=S // 5)
    batch = synth.simulate(stree, C=4, length=200, seed=5, gap_frac=0.01, missing_frac=0.03, ambig_frac=0.002)
    cfg = synth.write_files(batch, str(tmp_path), prefix="sim", params=dict(BURNIN=5, MCMC=10, CHAIN=1, GAPCHAR="-"))
    eng = pab.Engine.from_config(cfg, device=0)
    out = eng.gbgc_run(list(range(4)), stree.targets, traces=True, num_burn=10, num_mcmc=20, seed=3)
    assert np.isfinite(out["loglik_max"][out["filtered"] == 0]).all()
    el, Z, B, sn, sc, gb, want = [], [], [], [], [], [], []
    for i in range(4):
        if out["filtered"][i]:
            continue
        for mdl in range(3):
            assert out["accept"][i, mdl, 1] > 0  # some (Z, B) updates were accepted
            for m in (0, 7, 19, 29):
                el.append(i); Z.append(np.maximum(out["trace_z"][i, mdl, m], 0)); B.append(out["trace_b"][i, mdl, m])
                sc.append(out["trace"][i, mdl, 2, m]); sn.append(out["trace"][i, mdl, 3, m]); gb.append(out["trace"][i, mdl, 4, m])
                want.append(out["trace"][i, mdl, 0, m])
    got = eng.gbgc_replay(el, np.asarray(Z), np.asarray(B), sn, sc, gb)
    want = np.asarray(want)
    assert len(want) >= 24 and np.all(np.abs(got - want) <= 1e-10 * np.abs(want)), np.abs(got / want - 1).max()
    eng.close()
