"""GPU parity tests for the replay path (P(t) build + fp64 pruning over compressed patterns), through the C ABI.
Tolerance (north_star): replayed log-likelihoods within 1e-10 relative of the reference / oracle."""
import os
import tempfile

import numpy as np
import pytest

import phyloacc_b200 as pab
from oracle import coracle, frontend as fe
from phyloacc_b200 import synth
from tests._util import CASES, Case

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module", params=CASES)
def case(request):
    c = Case(request.param)
    yield c
    c.close()


def test_replay_matches_reference_golden(case, native_lib):
    """CUDA vs the log-likelihoods computed by the unmodified reference code (tests/golden/*/probe.txt)."""
    ref = case.golden_probe()
    reqs = [r for r in case.golden_requests() if "Z" in r]
    eng = pab.Engine.from_config(case.cfg, device=0)
    want, elem, Z, nr, cr = [], [], [], [], []
    k = 0
    for e in ref["elements"]:
        if e["filter"]:
            continue
        for st in e["states"]:
            r = reqs[k]
            k += 1
            want.append(st["loglik"]); elem.append(r["c"]); Z.append(r["Z"]); nr.append(r["n_rate"]); cr.append(r["c_rate"])
    got = eng.loglik_replay(elem, np.array(Z, dtype=np.int8), nr, cr)
    want = np.array(want)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() < TOL, rel.max()
    eng.close()


def _random_states(rng, el, N, model_resz, dollo=False):
    fixZ = fe.fixz_for_model(el, N, model_resz)
    Z = -np.ones(N, dtype=np.int8)
    for nd in reversed(el.nodes):
        if nd == el.root:
            Z[nd] = rng.integers(0, 2)
            continue
        zp = Z[el.parent2[nd]]
        opts = [0, 1] if zp == 0 else ([1] if fixZ[nd] == 1 else [1, 2])
        Z[nd] = opts[rng.integers(0, len(opts))]
    return Z


def _synth_engine(tree, batch, tmp, params=None):
    cfg = synth.write_files(batch, tmp, prefix="sim", params=params or dict(BURNIN=10, MCMC=20, CHAIN=1))
    P = fe.load_params(cfg)
    otree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    _, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(otree, P)
    X = fe.reorder_alignment(otree, names, X)
    return cfg, P, otree, model, X, starts, lens


def test_replay_matches_oracle_ragged_lengths(native_lib, oracle_lib):
    """cfg2-shaped ragged element lengths (50..2000 bp), gaps, ambiguity codes, missing species."""
    rng = np.random.default_rng(5)
    tree = synth.make_tree()
    lengths = [50, 51, 64, 127, 128, 129, 200, 333, 640, 1000, 2000, 57]
    batch = synth.simulate(tree, lengths=lengths, seed=3, gap_frac=0.02, missing_frac=0.05, ambig_frac=0.003)
    with tempfile.TemporaryDirectory() as tmp:
        cfg, P, otree, model, X, starts, lens = _synth_engine(tree, batch, tmp)
        eng = pab.Engine.from_config(cfg, device=0)
        elem, Zs, nr, cr, want = [], [], [], [], []
        for c in range(len(lengths)):
            el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
            if el.filter:
                continue
            h = coracle.ElementHandle(model, el)
            for resZ in (0, 2, 1):
                Z = _random_states(rng, el, otree.N, resZ)
                n, cc = float(rng.gamma(10, 0.2)), float(rng.gamma(5, 0.04))
                v, _ = h.loglik(resZ, Z, n, cc)
                elem.append(c); Zs.append(Z); nr.append(n); cr.append(cc); want.append(v)
        got = eng.loglik_replay(elem, np.array(Zs), nr, cr)
        want = np.array(want)
        assert (np.abs(got - want) / np.abs(want)).max() < TOL
        # filtered / empty inputs
        assert eng.loglik_replay([], np.zeros((0, otree.N), np.int8), [], []).shape == (0,)
        eng.close()


def test_replay_deep_tree_rescaling(native_lib, oracle_lib):
    """240-taxon tree with long branches: the 2^256 rescaling path of the pruning kernel must agree with log-space."""
    rng = np.random.default_rng(9)
    tree = synth.make_tree(S=240, seed=240, n_out=20, n_target=30, mean_bl=0.4)
    tree.dist[:] = np.clip(tree.dist * 6.0, 0, 3.0).astype(np.float32)  # saturate: site likelihoods ~ 0.25^240
    batch = synth.simulate(tree, C=3, length=60, seed=4, gap_frac=0.0, missing_frac=0.0)
    # rebuild the newick with the stretched branch lengths through the writer
    with tempfile.TemporaryDirectory() as tmp:
        import re
        it = iter(range(10 ** 9))
        tree2 = tree
        def fmt(n):
            if tree2.children[n, 0] < 0:
                return f"{tree2.names[n]}:{tree2.dist[n]:.6f}"
            a, b = tree2.children[n]
            return f"({fmt(a)},{fmt(b)}){tree2.names[n]}" + ("" if n == tree2.N - 1 else f":{tree2.dist[n]:.6f}")
        import sys
        sys.setrecursionlimit(10000)
        tree2.newick = fmt(tree2.N - 1) + ";"
        cfg, P, otree, model, X, starts, lens = _synth_engine(tree2, batch, tmp)
        eng = pab.Engine.from_config(cfg, device=0)
        elem, Zs, nr, cr, want = [], [], [], [], []
        for c in range(3):
            el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
            h = coracle.ElementHandle(model, el)
            Z = _random_states(rng, el, otree.N, 1)
            v, _ = h.loglik(1, Z, 3.0, 1.0)
            elem.append(c); Zs.append(Z); nr.append(3.0); cr.append(1.0); want.append(v)
        got = eng.loglik_replay(elem, np.array(Zs), nr, cr)
        want = np.array(want)
        assert np.isfinite(got).all()
        assert (np.abs(got - want) / np.abs(want)).max() < TOL
        eng.close()


def test_element_image_on_the_device_equals_the_host_image(native_lib):
    """pa_engine_element_blob: the image uploaded to HBM is the one the device-less engine keeps on the host."""
    from phyloacc_b200 import synth

    tree = synth.make_tree(S=10, seed=2, n_out=2, n_target=2)
    batch = synth.simulate(tree, C=4, length=100, seed=3, lengths=[60, 280, 90, 40], gap_frac=0.03, missing_frac=0.02, ambig_frac=0.01)
    prm = pab.default_params()
    engs = [pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                              conserve=tree.conserve, device=d) for d in (-1, 0)]
    for e in engs:
        e.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    n = 0
    for c in range(batch.C):
        host, dev = engs[0].element_blob(c), engs[1].element_blob(c)
        assert host.size == dev.size and np.array_equal(host, dev)
        assert (host.size == 0) == bool(engs[0].element_info(c, gg_max=400)["filtered"])
        n += host.size > 0
    assert n >= 3
    for e in engs:
        e.close()
