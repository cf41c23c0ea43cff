"""PhyloAcc-GT groundwork.  tests/golden/gt12 holds dumps of the UNMODIFIED reference (oracle/_ref/gt_probe, generator
tests/golden/make_gt_golden.py): 4 elements x 6 states -- gene trees with the species topology and draws from the multispecies coalescent,
random Z and rates -- each with the embedded gene tree as GTree holds it, the coalescent prior (GTree::priorTree, genetree.cpp:560-585),
log P(missing | Z) (get_logpZ1, :1156-1190) and the log-likelihood of the upward pass (Update_Tg, :1330-1418).
  * CPU: oracle/gt.py (numpy restatement; one pruning pass with EFFECTIVE branch lengths) against the dumps: prior and missing-data term
    to 1e-13, log-likelihood to 1e-12 relative -- which also shows that the ST integer front-end serves the GT variant on these inputs;
  * GPU: pa_engine_gt_replay (csrc/pa_gt.cu) against the same dumps to 1e-10 (north_star's replay criterion)."""
import os

import numpy as np
import pytest

from oracle import gt
from tests._util import Case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gt12")


def _golden():
    with open(os.path.join(GOLD, "probe.txt")) as f:
        return gt.parse_probe(f.read())


def test_species_heights_and_thetas():
    hdr, _ = _golden()
    case = Case("gt12")
    # the GT parser keeps branch lengths as doubles, PhyloAcc-ST rounds them through float (newick.cpp): 1e-8 apart, not bit-equal
    assert np.abs(gt.species_heights(case.tree.parent, case.tree.dist, hdr["N"]) - hdr["heights"]).max() < 1e-7
    assert np.allclose(hdr["thetas"][hdr["S"]:], 0.01, rtol=1e-4) and (hdr["thetas"][:hdr["S"]] == 0).all()
    case.close()


def test_restatement_matches_reference():
    hdr, els = _golden()
    case = Case("gt12")
    n = 0
    for e in els:
        assert not e["filter"]
        el = case.element(e["c"])
        assert el.GG == e["GG"]
        for st in e["states"]:
            assert abs(gt.prior_tree(st, hdr) - st["prior"]) <= 1e-13 * abs(st["prior"])
            assert abs(gt.logp_missing(st, hdr) - st["logpz"]) <= 1e-13 * abs(st["logpz"])
            assert np.array_equal(st["gmissing"][:hdr["S"]], el.missing[:hdr["S"]])
            ll = gt.replay_loglik(st, hdr, el.masks)
            assert abs(ll - st["loglik"]) <= 1e-12 * abs(st["loglik"]), (e["c"], ll, st["loglik"])
            n += 1
    assert n == 24 and any(st["coal"] for e in els for st in e["states"]) and any(st["gmissing"].any() for e in els for st in e["states"])
    case.close()


def test_coalescent_draws_differ_from_the_species_topology():
    """The fixture is only worth something if some gene trees are NOT the species tree."""
    hdr, els = _golden()
    differ = 0
    for e in els:
        ref = e["states"][0]["gchild"]
        differ += sum(1 for st in e["states"] if st["coal"] and not np.array_equal(np.sort(st["gchild"], axis=1), np.sort(ref, axis=1)))
    assert differ >= 4


def test_gt_replay_has_no_cpu_fallback_and_checks_its_arguments(native_lib):
    """Front-end-only engine (device = -1): the entry point refuses to compute (PA_ERR_CUDA, "no CPU fallback"); null arrays are PA_ERR_ARG."""
    import ctypes as C

    import phyloacc_b200 as pab

    hdr, els = _golden()
    case = Case("gt12")
    eng = pab.Engine.from_config(case.cfg, device=-1)
    st = dict(els[0]["states"][0], elem=0)
    with pytest.raises(pab.PhyloAccError) as ei:
        eng.gt_replay([st], hdr["heights"], hdr["thetas"])
    assert "no CPU fallback" in str(ei.value)
    gs = pab.PaGtStates()
    gs.n_states, gs.N = 1, hdr["N"]
    out = (C.c_double * 1)()
    assert pab.load_library().pa_engine_gt_replay(eng._h, C.byref(gs), out, None, None) == -1  # PA_ERR_ARG: null arrays
    eng.close()
    case.close()


@pytest.mark.gpu
def test_cuda_gt_replay_matches_reference(native_lib):
    import phyloacc_b200 as pab

    hdr, els = _golden()
    case = Case("gt12")
    eng = pab.Engine.from_config(case.cfg, device=0)
    states, want = [], []
    for e in els:
        for st in e["states"]:
            states.append(dict(st, elem=e["c"]))
            want.append((st["loglik"], st["prior"], st["logpz"]))
    want = np.array(want)
    ll, prior, lpz = eng.gt_replay(states, hdr["heights"], hdr["thetas"], cons_to_mis=states[0]["cons_to_mis"], ncons_to_mis=states[0]["ncons_to_mis"])
    assert np.all(np.abs(ll - want[:, 0]) <= 1e-10 * np.abs(want[:, 0])), np.abs(ll / want[:, 0] - 1).max()
    assert np.all(np.abs(prior - want[:, 1]) <= 1e-12 * np.abs(want[:, 1]))
    assert np.all(np.abs(lpz - want[:, 2]) <= 1e-12 * np.abs(want[:, 2]))
    # a second call in another order gives the same numbers (the per-CTA scratch carries nothing over)
    order = np.random.default_rng(0).permutation(len(states))
    ll2, _, _ = eng.gt_replay([states[i] for i in order], hdr["heights"], hdr["thetas"], cons_to_mis=states[0]["cons_to_mis"],
                              ncons_to_mis=states[0]["ncons_to_mis"])
    assert np.array_equal(ll2, ll[order])
    eng.close()
    case.close()


@pytest.mark.gpu
def test_gt_replay_on_the_species_tree_equals_the_st_replay(native_lib, tmp_path):
    """Two independent kernels, one property: a gene tree that IS the species tree (every coalescence exactly at its species node) has the
    species branch lengths as effective lengths, so pa_engine_gt_replay must give what pa_engine_loglik_replay gives for the same Z and
    rates -- 43 taxa, elements with gaps and ambiguity codes but no missing species (with missing species the ST path re-roots the tree,
    which is the same likelihood only as far as pi is stationary for the printed Q: 1e-6)."""
    import phyloacc_b200 as pab
    from oracle import frontend as fe
    from phyloacc_b200 import synth

    stree = synth.make_tree()
    batch = synth.simulate(stree, C=8, length=150, seed=3, gap_frac=0.03, missing_frac=0.0, ambig_frac=0.01)
    cfg = synth.write_files(batch, str(tmp_path), prefix="sim", params=dict(BURNIN=5, MCMC=10, CHAIN=1, SEED=1))
    eng = pab.Engine.from_config(cfg, device=0)
    tree = fe.parse_mod(fe.load_params(cfg).PHYTREE_FILE)
    N, S = tree.N, tree.S
    heights = gt.species_heights(np.asarray(tree.parent), np.asarray(tree.dist), N)
    children = np.asarray(tree.children).reshape(N, 2)
    cross = {s: [(s, s)] for s in range(S)}
    for s in range(S, N):
        cross[s] = [(int(children[s, 0]), s), (int(children[s, 1]), s), (s, s if s < N - 1 else N)]
    rng = np.random.default_rng(7)
    info = eng.run(list(range(batch.C)))
    states, elem, Zs, nr, cr = [], [], [], [], []
    for c in range(batch.C):
        if info.filtered[c]:
            continue
        for _ in range(3):
            Z = rng.integers(0, 3, N).astype(np.int8)
            n_rate, c_rate = 0.5 + 2 * rng.random(), 0.1 + 0.8 * rng.random()
            states.append(dict(elem=c, gchild=children, root=N - 1, gmissing=np.zeros(N, bool), gheights=heights, Z=Z, n_rate=n_rate, c_rate=c_rate,
                               cross=cross))
            elem.append(c); Zs.append(Z); nr.append(n_rate); cr.append(c_rate)
    assert len(states) >= 12
    ll_gt, _, _ = eng.gt_replay(states, heights, np.full(N, 0.01))
    ll_st = eng.loglik_replay(elem, np.asarray(Zs), nr, cr)
    assert np.all(np.isfinite(ll_st)) and np.all(np.abs(ll_gt - ll_st) <= 1e-10 * np.abs(ll_st)), np.abs(ll_gt / ll_st - 1).max()
    eng.close()


def test_emissions_given_a_sampled_history_match_reference():
    """tests/golden/gt12/probe_hist.txt: per element one species-topology and one coalescent gene tree, each with ONE substitution history
    drawn by the reference (GTree::Update_Tg(.., true, ..)) and the Z emissions BPP_C::getEmission (bpp_c.cpp:1043-1216) derives from it.  The
    emissions are a deterministic function of the dumped history: oracle/gt.py:emission restates them (loss terms, entry / exit states of
    every lineage of a species branch, the coalescences inside the branch summed out) -- 1e-12 on all three classes of every species."""
    with open(os.path.join(GOLD, "probe_hist.txt")) as f:
        hdr, els = gt.parse_probe(f.read())
    case = Case("gt12")
    n = 0
    for e in els:
        el = case.element(e["c"])
        for st in e["states"]:
            got = gt.emission(st, hdr, el.GG)
            assert set(st["emission"]) <= set(got)
            for s, want in st["emission"].items():
                assert np.all(np.abs(got[s] - want) <= 1e-12 * np.maximum(1.0, np.abs(want))), (e["c"], st["coal"], s, got[s], want)
                n += 1
    assert n >= 8 * 20
    case.close()
