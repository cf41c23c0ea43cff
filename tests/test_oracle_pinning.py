"""CPU tests: the oracle restatement is pinned against dumps of the UNMODIFIED reference code (tests/golden/*/probe.txt,
produced by tests/golden/make_golden.py with oracle/_ref/ref_probe), and the product's host-side integer front-end is
checked bit-exactly against the oracle through the C ABI (no CUDA involved)."""
import numpy as np
import pytest

import phyloacc_b200 as pab
from oracle import coracle, frontend as fe
from tests._util import CASES, Case, rel_err


@pytest.fixture(scope="module", params=CASES)
def case(request):
    c = Case(request.param)
    yield c
    c.close()


def test_tree_and_model_match_reference(case):
    ref = case.golden_probe()
    t = case.tree
    assert ref["S"] == t.S and ref["N"] == t.N
    assert (ref["parent"] == t.parent).all() and (ref["children"] == t.children).all()
    assert (ref["distances"] == t.dist).all()  # bit-exact, including the (float) rounding of newick.cpp:185
    assert (ref["pi"] == t.pi).all() and (ref["Q"] == t.Q).all()
    assert ref["names"] == t.names
    assert ref["upper"] == case.model.upper and ref["upper_conserve"] == case.model.upper_conserve
    assert (ref["element_size"] == case.lens).all()


def test_frontend_matches_reference_bit_exact(case):
    ref = case.golden_probe()
    assert len(ref["elements"]) == case.C
    n_pruned = 0
    for e in ref["elements"]:
        o = case.element(e["c"])
        assert o.filter == e["filter"]
        if o.filter:
            continue
        assert o.GG == e["GG"] and o.root == e["root"] and o.nodes == e["nodes"]
        assert (o.masks == e["masks"]).all() and (o.tg == e["tg"]).all()
        assert (o.missing == e["missing"]).all()
        for n in o.nodes:
            assert (o.children2[n] == e["children2"][n]).all()
            assert o.distances2[n] == e["distances2"][n]
            if n != o.root:
                assert o.parent2[n] == e["parent2"][n]
        assert o.upper_c == e["upper_c"] and o.upper_conserve_c == e["upper_conserve_c"]
        n_pruned += len(o.nodes) < case.tree.N
    if case.name == "ratite43":
        assert n_pruned > 0  # the fixture exercises pruned / re-rooted trees


def test_pattern_compression_roundtrip(case):
    for c in range(case.C):
        o = case.element(c)
        if o.filter:
            continue
        pt = fe.compress_patterns(o.masks)
        assert (pt.codes[pt.site_pattern] == o.masks).all()  # expanding patterns reproduces the reference's per-column encoding
        assert pt.weights.sum() == o.GG and (np.bincount(pt.site_pattern, minlength=pt.P) == pt.weights).all()
        assert (np.diff(pt.site_pattern[pt.site_order]) >= 0).all()
        rows = [tuple(r) for r in pt.codes]
        assert rows == sorted(set(rows))


def test_null_transition_matrices_match_reference(case, oracle_lib):
    ref = case.golden_probe()
    c0 = next(c for c in range(case.C) if not case.element(c).filter)
    h = case.handle(c0)
    for s, m in ref["logTMnull"].items():
        if case.tree.dist[s] > 1e-10:
            assert np.max(np.abs(h.log_tm(case.tree.dist[s], 1.0) - m)) < 1e-11


def test_replayed_loglik_matches_reference(case, oracle_lib):
    """north_star: replayed per-element log-likelihoods for (tree, Q, Z, r) states match to 1e-10 relative."""
    ref = case.golden_probe()
    reqs = [r for r in case.golden_requests() if "Z" in r]
    k = 0
    for e in ref["elements"]:
        if e["filter"]:
            continue
        h = case.handle(e["c"])
        for st in e["states"]:
            r = reqs[k]
            k += 1
            assert r["c"] == e["c"] and r["resZ"] == st["resZ"]
            v, per = h.loglik(r["resZ"], np.array(r["Z"], dtype=np.int8), r["n_rate"], r["c_rate"])
            assert abs(v - st["loglik"]) <= 1e-12 * abs(st["loglik"])
            assert rel_err(per, st["persite"]) < 1e-10
            assert ((st["fixZ"] == 1) == (fe.fixz_for_model(case.element(e["c"]), case.tree.N, r["resZ"]) == 1)).all()
    assert k == len(reqs) and k > 0


def test_emission_and_z_messages_match_reference(case, oracle_lib):
    ref = case.golden_probe()
    reqs = [r for r in case.golden_requests() if "history" in r]
    k = 0
    N, S = case.tree.N, case.tree.S
    for e in ref["elements"]:
        if e["filter"]:
            continue
        h = case.handle(e["c"])
        el = case.element(e["c"])
        for hist in e["histories"]:
            r = reqs[k]
            k += 1
            assert r["c"] == e["c"]
            tg = -np.ones((el.GG, N), dtype=np.int8)
            tg[:, S:] = np.array(r["history"]["hist"], dtype=np.int8)
            hh = r["history"]
            em, bk = h.emission_back(r["resZ"], np.array(r["Z"], dtype=np.int8), r["n_rate"], r["c_rate"], hh["g"], hh["l"], hh["l2"], tg)
            for s in el.nodes:
                for got, want in ((em[s], hist["emission"][s]), (bk[s], hist["back"][s])):
                    for a, b in zip(got, want):
                        if np.isinf(b) or np.isnan(b):
                            assert (np.isinf(a) and a == b) or (np.isnan(a) and np.isnan(b))
                        else:
                            assert abs(a - b) <= 1e-10 * max(1.0, abs(b))
    assert k == len(reqs) and k > 0


def test_eval2_matches_reference(case, oracle_lib):
    ref = case.golden_probe()
    reqs = [r for r in case.golden_requests() if "eval2" in r]
    by_c = {}
    for e in ref["elements"]:
        by_c[e["c"]] = e
    seen = {}
    n = 0
    for r in reqs:
        e = by_c[r["c"]]
        i = seen.get(r["c"], 0)
        seen[r["c"]] = i + 1
        want = e["eval2"][i]
        assert want["resZ"] == r["eval2"]
        h = case.handle(r["c"])
        tr, bufs = h.new_trace()
        t = r["trace"]
        for key in ("loglik", "n_rate", "c_rate", "g_rate", "l_rate", "l2_rate"):
            bufs[key][:] = t[key]
        bufs["Z"][:] = np.array(t["Z"], dtype=np.int8)
        # MonitorChain bookkeeping (bpp_c.cpp:768-814): full log-lik and arg-max after burn-in
        full = bufs["loglik"] + coracle.full_loglik_terms(case.P, r["eval2"], bufs["c_rate"], bufs["n_rate"])
        bufs["full_loglik"][:] = full
        post = np.arange(case.P.BURNIN, len(full))
        best = post[np.argmax(full[post])]  # first maximum, as the reference's strict '>' keeps the earliest
        bufs["max_Z"][:] = bufs["Z"][best]
        assert (bufs["max_Z"] == want["maxZ"]).all()
        assert abs(full[best] - want["maxloglik"]) <= 1e-10 * abs(want["maxloglik"])
        bufs["_trace"] = tr
        v = h.eval2(r["eval2"], bufs)
        if np.isnan(want["logml"]):
            assert np.isnan(v)
        elif np.isinf(want["logml"]):
            assert v == want["logml"]
        else:
            assert abs(v - want["logml"]) <= 1e-9 * abs(want["logml"]), (r["c"], r["eval2"], v, want["logml"])
        n += 1
    assert n > 0


def test_product_frontend_matches_oracle_through_c_abi(case, native_lib):
    eng = pab.Engine.from_config(case.cfg, device=-1)
    t = case.tree
    assert eng.S == t.S and eng.N == t.N
    assert (eng.parent == t.parent).all() and (eng.children == t.children).all() and (eng.distances == t.dist).all()
    assert (eng.pi == t.pi).all() and (eng.Q == t.Q).all()
    assert eng.node_names() == t.names
    assert list(eng.ids()) == fe.load_ids(case.P.ID_FILE, case.C)
    assert eng.num_elements == case.C
    for c in range(case.C):
        o = case.element(c)
        i = eng.element_info(c)
        assert i["filtered"] == o.filter
        if o.filter:
            continue
        pt = fe.compress_patterns(o.masks)
        assert i["GG"] == o.GG and i["P"] == pt.P and i["root"] == o.root and i["nodes"] == o.nodes
        assert (i["missing"] == o.missing).all()
        for n in o.nodes:
            assert (i["children2"][n] == o.children2[n]).all() and i["distances2"][n] == o.distances2[n]
            if n != o.root:
                assert i["parent2"][n] == o.parent2[n]
        assert (i["pattern_codes"] == pt.codes).all() and (i["pattern_weights"] == pt.weights).all()
        assert (i["site_pattern"] == pt.site_pattern).all()
        for m, resZ in pab.MODEL_RESZ.items():
            assert ((i["fixZ"][m] == 1) == (fe.fixz_for_model(o, t.N, resZ) == 1)).all()
    eng.close()


def test_c_abi_exports_every_declared_symbol(native_lib):
    import re, os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "phyloacc_b200.h")).read()
    declared = set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pab.EXPORTS)
    for name in declared:
        assert hasattr(native_lib, name), name


def test_compute_without_gpu_fails_loudly(case, native_lib):
    eng = pab.Engine.from_config(case.cfg, device=-1)
    with pytest.raises(pab.PhyloAccError):
        eng.loglik_replay([0], [np.ones(case.tree.N)], [1.0], [0.5])
    with pytest.raises(pab.PhyloAccError):
        eng.run([0])
    eng.close()
