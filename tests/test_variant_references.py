"""Groundwork for SURVEY.md §8 f1 / f2: the UNMODIFIED sources of PhyloAcc-ST-GBGC and PhyloAcc-GT build against the oracle's
Armadillo / GSL stand-ins (oracle/Makefile: ref-gbgc, ref-gt) and run end to end on a synthetic case, writing the files the
survey lists for them (GBGC/SRC/main.cpp:300-320: gBC + <node>_B columns; GT bpp.cpp:511-585, main.cpp:468-480: loglik_*_W,
_tree_M*.txt, _Beta_Post_pi_mode.txt).  These binaries are the future chain-level oracles and CPU baselines of the two variants;
no engine code is involved.  Skipped where the binaries are absent (they need /root/reference to build)."""
import os
import re
import subprocess

import pytest

from phyloacc_b200 import synth

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _case(tmp_path, extra):
    tree = synth.make_tree(S=8, seed=5, n_out=2, n_target=2)
    batch = synth.simulate(tree, C=2, length=80, seed=11, gap_frac=0.0, missing_frac=0.0)
    cfg = synth.write_files(batch, str(tmp_path), prefix="sim", params=dict(BURNIN=15, MCMC=30, CHAIN=1, SEED=2, NUM_THREAD=1, GAPCHAR="-"))
    with open(cfg, "a") as f:
        f.write(extra(tree, tmp_path))
    return tree, cfg


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "PhyloAcc-GBGC-ref")), reason="oracle/_ref/PhyloAcc-GBGC-ref not built")
def test_gbgc_reference_runs_on_the_shim(tmp_path):
    tree, cfg = _case(tmp_path, lambda t, d: "ADAPT_FREQ 10\n")
    r = subprocess.run([os.path.join(REF, "PhyloAcc-GBGC-ref"), cfg], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "time used:" in r.stdout
    out = tmp_path / "out"
    head = (out / "sim_rate_postZ_1.txt").read_text().split("\n")[0].split("\t")
    assert head[:4] == ["No.", "n_rate", "c_rate", "gBC"] and f"{tree.names[0]}_B" in head and f"{tree.names[0]}_3" in head
    rows = (out / "sim_elem_lik.txt").read_text().strip().split("\n")
    assert len(rows) == 3 and all(re.match(r"^-?[0-9.einf]+$", x) for x in rows[1].split("\t")[2:5])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "PhyloAcc-GT-ref")), reason="oracle/_ref/PhyloAcc-GT-ref not built")
def test_gt_reference_runs_on_the_shim(tmp_path):
    def extra(tree, d):
        theta = 0.01  # species tree in coalescent units = substitution lengths / (theta / 2) (PhyloAcc-GT main.cpp:376-400)
        coal = re.sub(r":([0-9.]+)", lambda m: ":%.6f" % (float(m.group(1)) / (theta / 2)), tree.newick)
        (d / "sim_coal.nwk").write_text(coal + "\n")
        return f"TREE_IN_COALESCENT_UNIT {d}/sim_coal.nwk\nTHIN 10\n"

    tree, cfg = _case(tmp_path, extra)
    r = subprocess.run([os.path.join(REF, "PhyloAcc-GT-ref"), cfg], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "time used:" in r.stdout
    out = tmp_path / "out"
    names = set(os.listdir(out))
    assert {"sim_elem_lik.txt", "sim_tree_M0.txt", "sim_tree_M1.txt", "sim_tree_M2.txt", "sim_M2_Beta_Post_pi_mode.txt", "sim_rate_postZ_M2.txt"} <= names
    head = (out / "sim_elem_lik.txt").read_text().split("\n")[0]
    assert "loglik_Null_W" in head and "loglik_Acc_W" in head and "loglik_Full_W" in head
