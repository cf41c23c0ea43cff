"""The PhyloAcc-ST-GBGC drop-in program (PhyloAcc-GBGC-b200 / pa_gbgc_run_from_config, csrc/pa_gbgc_app.cpp) against the UNMODIFIED
reference binary (oracle/_ref/PhyloAcc-GBGC-ref, src/PhyloAcc-ST-GBGC/SRC/main.cpp) on the same parameter file: same exit codes on the
argument / output-folder checks (CPU), and on a GPU the same set of result files with byte-identical headers, the same rows (element
numbers, ids, zero rows of elements that were not computed), the same column counts and number formats.  The chains themselves are compared
in tests/test_gbgc_chain.py; random streams differ, so values are checked for type and range here, not for equality."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from tests._util import Case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-GBGC-ref")
NUM = re.compile(r"^-?(\d+(\.\d*)?|\.\d+)(e[-+]?\d+)?$|^-?(inf|nan)$")


def _gbgc_cfg(case, out, ids, extra=""):
    """The variant's own keys (GBGC/SRC/main.cpp:108-193): the ST keys of the fixture that the variant does not know are simply ignored by it."""
    os.makedirs(out, exist_ok=True)
    idf = os.path.join(out, "ids.txt")
    with open(idf, "w") as f:
        f.write("".join(f"{i}\n" for i in ids))
    with open(case.cfg) as f:
        lines = [l for l in f.read().split("\n") if l and l.split()[0] not in ("ID_FILE", "RESULT_FOLDER", "BURNIN", "MCMC", "VERBOSE", "SEED")]
    cfg = os.path.join(out, "gbgc.cfg")
    with open(cfg, "w") as f:
        f.write("\n".join(lines) + f"\nID_FILE {idf}\nRESULT_FOLDER {out}\nBURNIN 20\nMCMC 60\nSEED 3\nCONSERVE_RATE -2\nPREP_GRATE 0.8\n{extra}")
    return cfg


def test_binary_probe_and_output_folder_check(native_lib):
    import phyloacc_b200 as pab

    assert os.path.exists(pab.BIN_GBGC_PATH)
    r = subprocess.run([pab.BIN_GBGC_PATH], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open the parameters file" in r.stderr  # the reference's no-argument behaviour (main.cpp:93-106)
    case = Case("gbgc_chain43")
    cfg = _gbgc_cfg(case, os.path.join(case.dir, "o1"), [0])
    with open(cfg) as f:
        body = f.read().replace(f"RESULT_FOLDER {case.dir}/o1", f"RESULT_FOLDER {case.dir}/does_not_exist")
    with open(cfg, "w") as f:
        f.write(body)
    ours = subprocess.run([pab.BIN_GBGC_PATH, cfg], capture_output=True, text=True)
    assert ours.returncode == 1 and "output path doesn't exist or empty!" in ours.stdout
    if os.path.exists(REF):
        ref = subprocess.run([REF, cfg], capture_output=True, text=True)
        assert ref.returncode == ours.returncode and "output path doesn't exist or empty!" in ref.stdout
    case.close()


def test_program_fails_loudly_without_a_gpu_and_on_bad_input(native_lib, tmp_path):
    """No CPU fallback in the GBGC program either: with a valid parameter file and no CUDA device it parses everything, then stops with exit
    code 1 and says why.  A parameter file that does not exist, SAMPLE_HYPER 1 and a chain too short for the estimator are refused before
    any data is read."""
    import phyloacc_b200 as pab
    import torch

    case = Case("gbgc_chain43")
    out = str(tmp_path / "o")
    cfg = _gbgc_cfg(case, out, [0, 1])
    r = subprocess.run([pab.BIN_GBGC_PATH, str(tmp_path / "missing.cfg")], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open the parameters file" in r.stderr
    if os.path.exists(REF):
        ref = subprocess.run([REF, str(tmp_path / "missing.cfg")], capture_output=True, text=True)
        assert ref.returncode == 1 and "Cannot open the parameters file" in ref.stderr
    for extra, msg in (("SAMPLE_HYPER 1\n", "SAMPLE_HYPER 1 is not supported"), ("MCMC 1\n", "MCMC >= 2")):
        with open(cfg) as f:
            body = f.read()
        bad = str(tmp_path / "bad.cfg")
        with open(bad, "w") as f:
            f.write(body + extra)
        r = subprocess.run([pab.BIN_GBGC_PATH, bad], capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (extra, r.stderr)
    if not torch.cuda.is_available():
        r = subprocess.run([pab.BIN_GBGC_PATH, cfg], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1
        assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
        assert not any(f.endswith("_elem_lik.txt") for f in os.listdir(out))
    case.close()


def _table(path):
    with open(path) as f:
        lines = f.read().split("\n")
    assert lines[-1] == ""  # every line is newline-terminated
    return [l.split("\t") for l in lines[:-1]]


@pytest.mark.gpu
def test_drop_in_writes_the_reference_files(native_lib):
    import phyloacc_b200 as pab

    if not os.path.exists(REF):
        pytest.skip("reference GBGC binary not built (make -C oracle ref-gbgc)")
    case = Case("gbgc_chain43")
    ids = [0, 3, 5, 7]
    outs = {}
    for tag, exe in (("ref", REF), ("ours", pab.BIN_GBGC_PATH)):
        out = os.path.join(case.dir, "o_" + tag)
        cfg = _gbgc_cfg(case, out, ids)
        r = subprocess.run([exe, cfg], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (tag, r.stdout[-2000:], r.stderr[-2000:])
        assert "time used:" in r.stdout
        outs[tag] = out
    skip = {"ids.txt", "gbgc.cfg"}
    files = {t: sorted(f for f in os.listdir(o) if f not in skip) for t, o in outs.items()}
    assert files["ours"] == files["ref"], (set(files["ours"]) ^ set(files["ref"]))
    N = case.tree.N
    for fn in files["ref"]:
        a, b = _table(os.path.join(outs["ref"], fn)), _table(os.path.join(outs["ours"], fn))
        if fn.endswith("_species_names.txt"):
            assert a == b
            continue
        assert a[0] == b[0], fn                                   # headers byte for byte
        assert len(a) == len(b), fn                               # same rows
        assert [len(r) for r in a] == [len(r) for r in b], fn     # same columns (trailing tabs included)
        if fn.endswith("_hyper.txt"):
            assert [r[0] for r in a] == [r[0] for r in b] and [r[1:5] for r in a[1:]] == [r[1:5] for r in b[1:]] and [r[7:] for r in a[1:]] == [r[7:] for r in b[1:]]
            continue
        for ra, rb in zip(a[1:], b[1:]):
            for x in rb:
                assert x == "" or NUM.match(x) or fn.endswith("_elem_lik.txt"), (fn, x)
        if "_rate_postZ_" in fn:
            assert [r[0] for r in a[1:]] == [r[0] for r in b[1:]] == [str(i) for i in ids]
            v = np.array([[float(x) for x in r] for r in b[1:]])
            pB, pZ = v[:, 4:4 + N], v[:, 4 + N:].reshape(len(ids), N, 4)
            assert ((pB >= 0) & (pB <= 1)).all() and ((pZ >= 0) & (pZ <= 1)).all()
            ref_v = np.array([[float(x) for x in r] for r in a[1:]])
            # nodes the front-end marks missing / pruned are the same on both sides: P(Z = -1) is 0 or 1 at the same places
            assert np.array_equal(ref_v[:, 4 + N:].reshape(len(ids), N, 4)[:, :, 0], pZ[:, :, 0])
            if fn.endswith("_0.txt"):  # the null model forbids the accelerated class
                assert (pZ[:, :, 3] == 0).all()
        elif fn.endswith("_elem_lik.txt"):
            assert [r[:2] for r in a[1:]] == [r[:2] for r in b[1:]]
            for ra, rb in zip(a[1:], b[1:]):
                computed = int(ra[0]) in ids
                assert (rb[2:] == ["0"] * 7) == (not computed) and (ra[2:] == ["0"] * 7) == (not computed)
                if computed:
                    va, vb = np.array([float(x) for x in ra[2:]]), np.array([float(x) for x in rb[2:]])
                    assert np.isfinite(vb).all() and np.all(np.abs(va - vb) < 0.05 * np.abs(va) + 25)  # 20 + 60 iterations: same scale only
        elif fn.endswith("_1_elem_Z.txt"):
            for k, (ra, rb) in enumerate(zip(a[1:], b[1:])):
                assert set(rb[:-1]) <= {"-1", "0", "1", "2"} and rb[-1] == ""
                if k not in ids:
                    assert ra == rb
                else:
                    assert [x == "-1" for x in ra] == [x == "-1" for x in rb]
        elif "_mcmc_trace_" in fn:
            v = np.array([[float(x) for x in r[:-1]] for r in b[1:]])
            assert v.shape == (80, 4 + 2 * N) and set(np.unique(v[:, 4:4 + N])) <= {0.0, 1.0} and set(np.unique(v[:, 4 + N:])) <= {-1.0, 0.0, 1.0, 2.0}
    case.close()
    shutil.rmtree(case.dir, ignore_errors=True)


@pytest.mark.gpu
def test_sharded_run_writes_byte_identical_files(native_lib):
    """pa_gbgc_run_from_config cuts the id list into shards (one engine + host thread each, pa_gbgc_app.cpp); chains are keyed by the element
    number, so every file -- VERBOSE traces included -- is byte-identical for one shard and for three shards on one GPU (and, on a box with
    several GPUs, for one shard per GPU)."""
    import phyloacc_b200 as pab
    import torch

    case = Case("gbgc_chain43")
    ids = list(range(case.C))
    runs = {"one": {"PA_DEVICES": "1"}, "three_on_gpu0": {"PA_DEVICE_LIST": "0,0,0", "PA_MIN_ELEMS_PER_DEVICE": "2"}}
    if torch.cuda.device_count() > 1:
        runs["per_gpu"] = {"PA_MIN_ELEMS_PER_DEVICE": "2"}
    outs = {}
    for tag, env in runs.items():
        out = os.path.join(case.dir, "s_" + tag)
        cfg = _gbgc_cfg(case, out, ids)
        r = subprocess.run([pab.BIN_GBGC_PATH, cfg], capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
        assert r.returncode == 0, (tag, r.stdout[-1500:], r.stderr[-1500:])
        if tag != "one":
            assert "# element shards =" in r.stdout
        outs[tag] = out
    names = sorted(f for f in os.listdir(outs["one"]) if f not in ("ids.txt", "gbgc.cfg"))
    assert len(names) > 3 * len(ids)
    for tag, out in outs.items():
        assert sorted(f for f in os.listdir(out) if f not in ("ids.txt", "gbgc.cfg")) == names
        for fn in names:
            with open(os.path.join(outs["one"], fn), "rb") as fa, open(os.path.join(out, fn), "rb") as fb:
                assert fa.read() == fb.read(), (tag, fn)
    case.close()


@pytest.mark.gpu
def test_batch_selection_and_outer_chain_iterations_follow_the_reference(native_lib):
    """No ID_FILE: BATCH k takes the k-th third of the BED file (main.cpp:331-347).  CHAIN 2: two outer iterations -- the _rate_postZ files
    get the rows of both, the VERBOSE traces are appended to, _hyper.txt gets one line per iteration, _elem_lik.txt is rewritten."""
    import phyloacc_b200 as pab

    if not os.path.exists(REF):
        pytest.skip("reference GBGC binary not built (make -C oracle ref-gbgc)")
    case = Case("gbgc_chain43")
    outs = {}
    for tag, exe in (("ref", REF), ("ours", pab.BIN_GBGC_PATH)):
        out = os.path.join(case.dir, "b_" + tag)
        cfg = _gbgc_cfg(case, out, [0])
        with open(cfg) as f:
            lines = [l for l in f.read().split("\n") if l and l.split()[0] not in ("ID_FILE", "CHAIN")]
        with open(cfg, "w") as f:
            f.write("\n".join(lines) + "\nBATCH 1\nCHAIN 2\n")
        r = subprocess.run([exe, cfg], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (tag, r.stdout[-1500:], r.stderr[-1500:])
        assert "Running MCMC chain 2 ..." in r.stdout
        outs[tag] = out
    skip = {"ids.txt", "gbgc.cfg"}
    files = {t: sorted(f for f in os.listdir(o) if f not in skip) for t, o in outs.items()}
    assert files["ours"] == files["ref"]
    third = case.C // 3
    want_ids = [str(c) for c in range(third, 2 * third)]
    for fn in files["ref"]:
        a, b = _table(os.path.join(outs["ref"], fn)), _table(os.path.join(outs["ours"], fn))
        assert a[0] == b[0] and [len(r) for r in a] == [len(r) for r in b], fn
        if "_rate_postZ_" in fn:
            assert [r[0] for r in a[1:]] == [r[0] for r in b[1:]] == want_ids + want_ids
        elif "_mcmc_trace_" in fn:
            assert len(b) == 1 + 2 * 80
        elif fn.endswith("_hyper.txt"):
            assert [r[0] for r in b[1:]] == ["0", "1"]
    case.close()
