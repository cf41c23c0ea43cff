"""GPU parity tests for the on-device chain (history sampling, Z Gibbs, MH rate updates, Beta updates) through the
C ABI.  The oracle (oracle/oracle.c, pinned against the reference) follows the same Philox draw schedule, so chains
are compared DRAW FOR DRAW: integer traces (Z) must be identical, floating-point traces within 1e-9 relative."""
import numpy as np
import pytest

import phyloacc_b200 as pab
from oracle import coracle, frontend as fe
from tests._util import CASES, Case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=CASES)
def case(request):
    c = Case(request.param)
    yield c
    c.close()


def _oracle_chain(case, c, mdl):
    h = case.handle(c, canonical_sites=True)
    tr = h.chain(pab.MODEL_RESZ[mdl], case.P.SEED, c * 4 + mdl)
    return h, tr


def test_chain_traces_match_oracle_draw_for_draw(case, native_lib, oracle_lib):
    eng = pab.Engine.from_config(case.cfg, device=0)
    ids = list(range(case.C))
    res = eng.run(ids)
    n_chains = 0
    for i, c in enumerate(ids):
        el = case.element(c)
        assert bool(res.filtered[i]) == el.filter
        if el.filter:
            continue
        for mdl in (pab.M0, pab.M1, pab.M2):
            h, otr = _oracle_chain(case, c, mdl)
            tr = eng.trace(i, mdl)
            assert (tr["Z"] == otr["Z"]).all(), (c, mdl, np.argwhere(tr["Z"] != otr["Z"])[:5])
            for k in ("loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate"):
                a, b = tr[k], otr[k]
                assert np.all(np.abs(a - b) <= 1e-9 * np.maximum(1e-300, np.abs(b))), (c, mdl, k, np.max(np.abs(a - b)))
            # summaries (MaxLoglik / Max_Z / Eval2 / medians / posterior Z) against the oracle's on its own trace
            assert abs(res.loglik_max[i, mdl] - otr["max_loglik"]) <= 1e-9 * abs(otr["max_loglik"])
            assert (res.max_Z[i, mdl] == otr["max_Z"]).all()
            # Eval2 (bpp_c2.cpp:14-353): the product's host implementation vs the oracle's, both fed the CUDA trace (the
            # log-det of a few-sample covariance is ill-conditioned, so 1e-9 differences between traces are not comparable)
            tro, bufs = h.new_trace()
            for k in ("loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate"):
                bufs[k][:] = tr[k]
            bufs["Z"][:] = tr["Z"]
            full = bufs["loglik"] + coracle.full_loglik_terms(case.P, pab.MODEL_RESZ[mdl], bufs["c_rate"], bufs["n_rate"])
            bufs["full_loglik"][:] = full
            post = np.arange(case.P.BURNIN, len(full))
            bufs["max_Z"][:] = bufs["Z"][post[np.argmax(full[post])]]
            bufs["_trace"] = tro
            want = h.eval2(pab.MODEL_RESZ[mdl], bufs)
            got = res.loglik[i, mdl]
            if np.isfinite(want):
                assert abs(got - want) <= 1e-9 * abs(want), (c, mdl, got, want)
            else:
                assert (np.isnan(want) and np.isnan(got)) or got == want
            T = len(otr["loglik"])
            med = [np.sort(tr[k])[T // 2] for k in ("n_rate", "c_rate", "g_rate", "l_rate", "l2_rate")]
            assert np.array_equal(res.median_rates[i, mdl], np.array(med))
            post = otr["Z"][case.P.BURNIN:]
            for s in range(case.tree.N):
                cnt = np.bincount(post[:, s] + 1, minlength=4).astype(float)
                if el.missing[s]:
                    cnt[0] = case.P.MCMC
                assert np.allclose(res.post_Z[i, mdl, s], cnt / case.P.MCMC)
            n_chains += 1
    assert n_chains >= 6
    cnt = eng.counters()
    assert cnt["kernel_launches"] == 2 and cnt["chain_iterations"] == n_chains * (case.P.BURNIN + case.P.MCMC - 1)
    eng.close()


def test_chain_is_deterministic_and_seed_dependent(native_lib):
    case = Case("tiny8")
    eng = pab.Engine.from_config(case.cfg, device=0)
    ids = [c for c in range(case.C) if not case.element(c).filter][:2]
    a = eng.run(ids)
    b = eng.run(ids)
    assert np.array_equal(a.loglik_max, b.loglik_max) and np.array_equal(a.post_Z, b.post_Z)
    eng.close()
    case2 = Case("tiny8", extra_params=dict(SEED=99))
    eng2 = pab.Engine.from_config(case2.cfg, device=0)
    c2 = eng2.run(ids)
    assert not np.array_equal(a.post_Z, c2.post_Z)
    eng2.close()
    case.close()
    case2.close()


def _synthetic_case(tmp, S, C, length, params, tree_seed=9, seed=4, **sim_kw):
    from phyloacc_b200 import synth

    tree = synth.make_tree(S=S, seed=tree_seed, n_out=max(2, S // 6), n_target=max(2, S // 5))
    batch = synth.simulate(tree, C=C, length=length, seed=seed, **sim_kw)
    cfg = synth.write_files(batch, tmp, prefix="syn", params=params)
    P = fe.load_params(cfg)
    otree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    _, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(otree, P)
    X = fe.reorder_alignment(otree, names, X)
    return cfg, P, otree, model, X, starts, lens


def _compare_with_oracle(eng, P, model, X, starts, lens, ids, tol=1e-9):
    res = eng.run(ids)
    n = 0
    for i, c in enumerate(ids):
        el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
        assert bool(res.filtered[i]) == el.filter
        if el.filter:
            continue
        h = coracle.ElementHandle(model, el, fe.compress_patterns(el.masks).site_order)
        for mdl in (pab.M0, pab.M1, pab.M2):
            otr = h.chain(pab.MODEL_RESZ[mdl], P.SEED, c * 4 + mdl)
            tr = eng.trace(i, mdl)
            assert (tr["Z"] == otr["Z"]).all(), (c, mdl, np.argwhere(tr["Z"] != otr["Z"])[:5])
            for k in ("loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate"):
                assert np.all(np.abs(tr[k] - otr[k]) <= tol * np.maximum(1e-300, np.abs(otr[k]))), (c, mdl, k)
            n += 1
    return n


def test_deep_tree_240_taxa_matches_oracle(native_lib, oracle_lib, tmp_path):
    """BASELINE config 3 shape (Zoonomia-like, 240 species = 479 nodes): the deep-tree layout of the chain kernel (one CTA
    per SM, tip codes read from the blob, more branches than threads in the pair-count pass); 2^256 rescaling is exercised."""
    cfg, P, otree, model, X, starts, lens = _synthetic_case(str(tmp_path), S=240, C=3, length=90, gap_frac=0.02, missing_frac=0.02,
                                                            params=dict(BURNIN=4, MCMC=10, CHAIN=1, ADAPT_FREQ=5, SEED=6))
    eng = pab.Engine.from_config(cfg, device=0)
    assert _compare_with_oracle(eng, P, model, X, starts, lens, [0, 1, 2]) >= 6
    # replayed likelihood of the initial state on the deep tree
    el = fe.element_frontend(model, X, int(starts[0]), int(lens[0]))
    Z = -np.ones(otree.N, dtype=np.int8)
    Z[el.nodes] = 1
    Z[el.root] = 0
    want, _ = coracle.ElementHandle(model, el).loglik(1, Z, 1.0, 0.5)
    got = eng.loglik_replay([0], Z[None], [1.0], [0.5])[0]
    assert abs(got - want) <= 1e-10 * abs(want)
    eng.close()


def test_balanced_tree_parks_deeper_than_the_shared_memory_stack(native_lib, oracle_lib, tmp_path, monkeypatch):
    """A perfectly balanced tree needs more temporaries than the four stack slots the 256-thread kernel keeps in shared
    memory: the deeper parks are re-read from the arena (build_pass_program), with the same chains as the oracle."""
    from phyloacc_b200 import synth

    def balanced(rng, n_tips):
        if n_tips == 1:
            return ("tip",)
        return (balanced(rng, n_tips // 2), balanced(rng, n_tips - n_tips // 2))

    monkeypatch.setattr(synth, "_yule_topology", balanced)
    cfg, P, otree, model, X, starts, lens = _synthetic_case(str(tmp_path), S=40, C=3, length=120, gap_frac=0.01, missing_frac=0.0,
                                                            params=dict(BURNIN=6, MCMC=14, CHAIN=1, ADAPT_FREQ=5, SEED=9))
    eng = pab.Engine.from_config(cfg, device=0)
    assert max(eng.element_info(c)["n_slots"] for c in range(3)) > 4
    assert _compare_with_oracle(eng, P, model, X, starts, lens, [0, 1, 2]) >= 6
    eng.close()


@pytest.mark.parametrize("S", [14, 110, 130])
def test_ambiguity_codes_and_several_site_tiles_match_oracle(native_lib, oracle_lib, tmp_path, S):
    """Elements longer than one 256-site tile with ambiguity-coded tips (r, y, k, m, s, w): the pair-count pass takes the tips from
    the per-tile bit planes of the element blob, ambiguous sites through the out-of-line path (bpp_c2.cpp:372-378); S = 130 runs on
    a deep-tree layout, where the sampled bases exist only as bit planes."""
    from phyloacc_b200 import synth

    tree = synth.make_tree(S=S, seed=11, n_out=max(2, S // 6), n_target=max(2, S // 5))
    batch = synth.simulate(tree, C=3, length=100, seed=8, lengths=[300, 521, 64], gap_frac=0.03, missing_frac=0.02, ambig_frac=0.02)
    cfg = synth.write_files(batch, str(tmp_path), prefix="amb", params=dict(BURNIN=3, MCMC=6, CHAIN=1, ADAPT_FREQ=4, SEED=12))
    P = fe.load_params(cfg)
    otree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    _, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(otree, P)
    X = fe.reorder_alignment(otree, names, X)
    assert np.isin(X, np.frombuffer(b"rykmsw", dtype=np.uint8)).sum() > 50
    eng = pab.Engine.from_config(cfg, device=0)
    assert _compare_with_oracle(eng, P, model, X, starts, lens, [0, 1, 2]) >= 6
    eng.close()
