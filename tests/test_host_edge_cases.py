"""Host side of the path (b): parsers, filters and the error behaviour of the drop-in binary on edge
inputs.  No GPU needed: every check stops before (or never reaches) a CUDA call.

Where the reference binary built from the unmodified sources is available (oracle/_ref/PhyloAcc-ST-ref, built by
`make -C oracle ref`; it travels to the GPU box) the drop-in binary is compared with it directly: same exit code, same
message.  Front-end results on edge inputs are compared with the oracle front-end (pinned bit-exact against the
reference on the golden cases, which include trimming, PRUNE_TREE, MIN_LEN and GAP_PROP)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import phyloacc_b200 as pab
from oracle import frontend as fe, probe
from phyloacc_b200 import synth
from tests._util import Case

REF_BIN = probe.REF_BIN if os.path.exists(probe.REF_BIN) else None


def _run(binary, *args):
    r = subprocess.run([binary, *args], capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout + r.stderr


def _both(*args):
    """(ours, reference or None)"""
    ours = _run(pab.BIN_PATH, *args)
    ref = _run(REF_BIN, *args) if REF_BIN else None
    return ours, ref


def test_no_argument_probe(native_lib):
    # the driver's presence probe (phyloacc_lib/core.py:47-110) runs the binary with no arguments and wants exit code <= 1
    (rc, out), ref = _both()
    assert rc == 1 and "Please specify the path of the parameter file." in out
    if ref:
        assert ref[0] == rc and "Please specify the path of the parameter file." in ref[1]


def test_unreadable_config(native_lib, tmp_path):
    (rc, out), ref = _both(str(tmp_path / "nope.cfg"))
    assert rc == 1 and "Cannot open the parameters file" in out
    if ref:
        assert ref[0] == 1 and "Cannot open the parameters file" in ref[1]


def test_missing_result_folder(native_lib):
    case = Case("tiny8")
    shutil.rmtree(os.path.join(case.dir, "out"))
    (rc, out), ref = _both(case.cfg)
    assert rc == 1 and "output path doesn't exist or empty!" in out  # main.cpp:334-338, checked before any data is loaded
    if ref:
        assert ref[0] == 1 and "output path doesn't exist or empty!" in ref[1]
    case.close()


@pytest.mark.parametrize("which,msg", [("sim.fa", "Cannot open the phylogenetic profile input file"), ("sim.bed", "Cannot open the segment input file"),
                                       ("sim.mod", "Cannot open the phylogenetic tree input file")])
def test_missing_input_files(which, msg, native_lib):
    case = Case("tiny8")
    os.remove(os.path.join(case.dir, which))
    (rc, out), ref = _both(case.cfg)
    assert rc == 1 and msg in out, out[-500:]  # profile.cpp:26-29, 74-77; newick.cpp:438-442
    if ref:
        assert ref[0] == 1 and msg in ref[1]
    # the C ABI reports the same condition as an error code + message, without exiting the host process
    with pytest.raises(pab.PhyloAccError) as ei:
        pab.Engine.from_config(case.cfg, device=-1)
    assert ei.value.code in (-4, -1) and msg in str(ei.value)
    case.close()


def test_unlabeled_internal_node(native_lib):
    case = Case("tiny8")
    mod = os.path.join(case.dir, "sim.mod")
    text = open(mod).read()
    # drop the label of the last internal node before the root's closing parenthesis
    import re
    tree_line = [l for l in text.split("\n") if l.startswith("TREE:")][0]
    broken = re.sub(r"\)([A-Za-z_][A-Za-z0-9_]*):", "):", tree_line, count=1)
    assert broken != tree_line
    open(mod, "w").write(text.replace(tree_line, broken))
    (rc, out), ref = _both(case.cfg)
    assert rc == 1 and "Some nodes in the tree aren't labeled" in out  # newick.cpp:416-419
    if ref:
        assert ref[0] == 1 and "Some nodes in the tree aren't labeled" in ref[1]
    case.close()


def test_config_quirks_unknown_keys_comments_blank_lines(native_lib):
    case = Case("tiny8")
    body = open(case.cfg).read()
    open(case.cfg, "w").write("# a comment line\n\n" + body + "\nNOT_A_KEY 12\n\n")
    eng = pab.Engine.from_config(case.cfg, device=-1)  # unknown keys are reported and skipped (main.cpp:280-283)
    assert eng.num_elements == case.C
    p = eng.params()
    assert p.burnin == case.P.BURNIN and p.mcmc == case.P.MCMC and p.seed == case.P.SEED and p.hyper_lrate2_a == case.P.HYPER_LRATE2_A
    eng.close()
    case.close()


def test_bed_stops_at_first_short_line(native_lib):
    # profile.cpp:84: the element list ends at the first line with fewer than three tab-separated fields
    case = Case("ratite43")
    bed = os.path.join(case.dir, "sim.bed")
    lines = open(bed).read().strip().split("\n")
    keep = 4
    open(bed, "w").write("\n".join(lines[:keep] + ["truncated"] + lines[keep:]) + "\n")
    open(os.path.join(case.dir, "sim.id"), "w").write("\n".join(str(i) for i in range(keep)) + "\n")
    eng = pab.Engine.from_config(case.cfg, device=-1)
    names, pos = fe.load_bed(bed)
    assert len(names) == keep and eng.num_elements == keep
    eng.close()
    case.close()


def _front_end_case(tmp, **kw):
    tree = synth.make_tree(S=12, seed=3, n_out=2, n_target=3)
    batch = synth.simulate(tree, C=kw.pop("C", 6), length=kw.pop("length", 80), seed=kw.pop("seed", 5), **kw.pop("sim", {}))
    cfg = synth.write_files(batch, tmp, prefix="edge", params=kw.pop("params", {}))
    P = fe.load_params(cfg)
    otree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    _, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(otree, P)
    return cfg, model, fe.reorder_alignment(otree, names, X), starts, lens


def _check_frontend(cfg, model, X, starts, lens):
    eng = pab.Engine.from_config(cfg, device=-1)
    n_f = 0
    for c in range(len(starts)):
        o = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
        i = eng.element_info(c)
        assert i["filtered"] == o.filter, c
        if o.filter:
            n_f += 1
            continue
        pt = fe.compress_patterns(o.masks)
        assert i["GG"] == o.GG and i["P"] == pt.P and i["root"] == o.root and i["nodes"] == o.nodes
        assert (i["missing"] == o.missing).all() and (i["pattern_codes"] == pt.codes).all() and (i["pattern_weights"] == pt.weights).all()
    eng.close()
    return n_f


def test_short_and_gappy_elements_are_filtered_like_the_reference(native_lib, tmp_path):
    # MIN_LEN filter (bpp_c.hpp:158), column trimming (bpp_c.hpp:268-301), missing-species / conserved filter (bpp_c.hpp:331-373)
    cfg, model, X, starts, lens = _front_end_case(str(tmp_path), length=70, sim=dict(gap_frac=0.25, missing_frac=0.3, ambig_frac=0.02),
                                                  params=dict(MIN_LEN=60, TRIM_GAP_PERCENT=0.5, GAP_PROP=0.4, CONSERVE_PROP=0.3))
    n_f = _check_frontend(cfg, model, X, starts, lens)
    assert 0 < n_f  # the case really exercises the filters


def test_prune_tree_and_gapchar(native_lib, tmp_path):
    cfg, model, X, starts, lens = _front_end_case(str(tmp_path), length=90, sim=dict(gap_frac=0.1, missing_frac=0.25), params=dict(PRUNE_TREE=1, GAP_PROP=0.5))
    _check_frontend(cfg, model, X, starts, lens)


def test_element_past_the_alignment_end_is_clipped(native_lib, tmp_path):
    # std::string::substr semantics (bpp_c.hpp:211): an element that runs past the alignment is shortened, not rejected
    cfg, model, X, starts, lens = _front_end_case(str(tmp_path), C=3, length=100)
    bed = fe.load_params(cfg).SEG_FILE
    rows = [l.split("\t") for l in open(bed).read().strip().split("\n")]
    rows[-1][2] = str(int(rows[-1][2]) + 40)  # the last element claims 40 more columns than the FASTA has
    open(bed, "w").write("\n".join("\t".join(r) for r in rows) + "\n")
    eng = pab.Engine.from_config(cfg, device=-1)
    i = eng.element_info(2)
    assert not i["filtered"] and i["GG"] == 100
    eng.close()


def test_empty_id_list_and_out_of_range_ids(native_lib, tmp_path):
    case = Case("tiny8")
    open(os.path.join(case.dir, "sim.id"), "w").write("")
    eng = pab.Engine.from_config(case.cfg, device=-1)
    assert len(eng.ids()) == 0
    eng.close()
    case.close()


def test_element_longer_than_the_device_index_range_is_refused(native_lib):
    # pattern / site indices are 16-bit on the device: a 70 kb "element" is an error (PA_ERR_LIMIT), not a silent wrap-around
    tree = synth.make_tree(S=8, seed=5, n_out=2, n_target=2)
    G = 70000
    X = np.full((tree.S, G), ord("a"), dtype=np.uint8)
    prm = pab.default_params()
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup, conserve=tree.conserve, device=-1)
    with pytest.raises(pab.PhyloAccError) as ei:
        eng.load_alignment(X, np.array([0], dtype=np.int64), np.array([G], dtype=np.int32))
    assert ei.value.code == -5 and "65535" in str(ei.value)
    eng.load_alignment(X, np.array([0], dtype=np.int64), np.array([60000], dtype=np.int32))  # below the limit: accepted
    assert not eng.element_info(0, gg_max=60000)["filtered"]
    eng.close()


def test_element_image_tip_planes_follow_codes_and_site_order(native_lib):
    """The packed element image (ElemHeader, csrc/pa_internal.h): the 4-bit codes per pattern, the site -> pattern map in the
    canonical site order (sorted by pattern, then column) and the tips' bit planes per 256-site tile that the chain kernel's
    pair-count pass intersects with the sampled bases -- bit t of word w of plane b of species s in tile k is set iff the code
    of s at site 256 k + 32 w + t is exactly base b; ambiguity codes go to the `tipamb` plane, gaps / N to none."""
    from phyloacc_b200 import synth

    tree = synth.make_tree(S=9, seed=4, n_out=2, n_target=2)
    batch = synth.simulate(tree, C=3, length=100, seed=6, lengths=[70, 300, 515], gap_frac=0.05, missing_frac=0.0, ambig_frac=0.03)
    prm = pab.default_params()
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                            conserve=tree.conserve, device=-1)
    eng.load_alignment(batch.X, batch.elem_start, batch.elem_len)
    names = ["n", "ni", "S", "P", "Ppad", "GG", "W", "n_slots", "n_up", "n_down", "elem_id", "blob_bytes", "off_parent", "off_child",
             "off_flags", "off_tiprow", "off_orig", "off_dist", "off_int", "off_intidx", "off_su", "off_slot", "off_up_nodes", "off_up_off",
             "off_down_nodes", "off_down_off", "off_codes", "off_weights", "off_gap", "off_sitepat", "off_tipplanes", "off_tipamb"]
    n_amb = 0
    for c in range(3):
        info = eng.element_info(c, gg_max=600)
        assert not info["filtered"]
        blob = eng.element_blob(c)
        h = dict(zip(names, np.frombuffer(blob[: 4 * len(names)].tobytes(), dtype=np.int32)))
        S, P, Ppad, GG = int(h["S"]), int(h["P"]), int(h["Ppad"]), int(h["GG"])
        assert h["blob_bytes"] == blob.size and S == tree.S and P == info["P"] and GG == info["GG"] and Ppad % 32 == 0
        packed = blob[h["off_codes"]: h["off_codes"] + S * (Ppad // 2)].reshape(S, Ppad // 2)
        codes = np.empty((S, Ppad), dtype=np.uint8)
        codes[:, 0::2] = packed & 15
        codes[:, 1::2] = packed >> 4
        assert np.array_equal(codes[:, :P].T, info["pattern_codes"])
        sitepat = np.frombuffer(blob[h["off_sitepat"]: h["off_sitepat"] + 2 * GG].tobytes(), dtype=np.uint16).astype(int)
        assert np.all(np.diff(sitepat) >= 0) and np.array_equal(np.bincount(sitepat, minlength=P), info["pattern_weights"])
        tiles = (GG + 255) // 256
        planes = np.frombuffer(blob[h["off_tipplanes"]: h["off_tipplanes"] + tiles * S * 8 * 16].tobytes(), dtype=np.uint32).reshape(tiles, S, 8, 4)
        amb = np.frombuffer(blob[h["off_tipamb"]: h["off_tipamb"] + tiles * S * 8 * 4].tobytes(), dtype=np.uint32).reshape(tiles, S, 8)
        site_codes = codes[:, sitepat]  # [S][GG] in canonical site order
        for s in range(S):
            for g in range(tiles * 256):
                k, w, t = g >> 8, (g >> 5) & 7, g & 31
                bits = [(int(planes[k, s, w, b]) >> t) & 1 for b in range(4)]
                a = (int(amb[k, s, w]) >> t) & 1
                code = int(site_codes[s, g]) if g < GG else 15
                if code in (1, 2, 4, 8):
                    assert bits == [int(code == (1 << b)) for b in range(4)] and a == 0
                elif code == 15:
                    assert bits == [0, 0, 0, 0] and a == 0
                else:
                    assert bits == [0, 0, 0, 0] and a == 1
                    n_amb += 1
    assert n_amb > 10
    eng.close()


def test_drop_in_binary_fails_loudly_without_a_gpu(native_lib, tmp_path):
    """There is no CPU fallback anywhere in the product: with a valid parameter file and no CUDA device the binary parses everything,
    then stops with exit code 1 and says why (the driver's log shows it); with no arguments it still answers the driver's probe."""
    import subprocess

    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phyloacc_b200 import synth

    tree = synth.make_tree(S=8, seed=5, n_out=2, n_target=2)
    batch = synth.simulate(tree, C=3, length=80, seed=11)
    cfg = synth.write_files(batch, str(tmp_path), prefix="sim", params=dict(BURNIN=5, MCMC=10))
    r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert not any(f.endswith("_elem_lik.txt") for f in os.listdir(tmp_path / "out"))
