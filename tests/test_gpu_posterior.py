"""GPU tests against runs of the UNMODIFIED reference binary (tests/golden/post43/refrun_seed*.json, produced by
oracle/_ref/PhyloAcc-ST-ref): chains use different random streams (mt19937 vs Philox), so agreement is
distributional -- marginal-likelihood estimates, Bayes factors, posterior Z probabilities and rate medians must agree
within Monte-Carlo error across seeds, and the accelerated-element call set must match at the post-processing
cut-offs (phyloacc_post.py:245-262, cut-off 5).  Also checks the drop-in binary's output files."""
import os
import subprocess

import numpy as np
import pytest

import phyloacc_b200 as pab
from tests._util import Case, parse_table

pytestmark = pytest.mark.gpu


def _ref_tables(case):
    runs = case.golden_refruns()
    lik = []
    post = {m: [] for m in range(3)}
    for r in runs:
        h, rows = parse_table(r["sim_elem_lik.txt"])
        lik.append(np.array([[float(x) for x in row[2:]] for row in rows]))
        for m in range(3):
            h2, rows2 = parse_table(r[f"sim_rate_postZ_M{m}.txt"])
            rows2 = sorted(rows2, key=lambda t: int(t[0]))
            post[m].append(np.array([[float(x) for x in row[1:]] for row in rows2]))
    return np.array(lik), {m: np.array(v) for m, v in post.items()}


def test_posteriors_agree_with_reference_runs(native_lib):
    base = Case("post43")
    ref_lik, ref_post = _ref_tables(base)  # [seeds, C, 8], [seeds, C, 5 + 4N]
    C, N = base.C, base.tree.N
    ours_lik, ours_med, ours_pz = [], [], []
    for seed in range(11, 23):
        case = Case("post43", extra_params=dict(SEED=seed))
        eng = pab.Engine.from_config(case.cfg, device=0)
        res = eng.run(list(range(C)))
        assert not res.filtered.any()
        ours_lik.append(res.loglik.copy())
        ours_med.append(res.median_rates.copy())
        ours_pz.append(res.post_Z.copy())
        eng.close()
        case.close()
    ours_lik, ours_med, ours_pz = np.array(ours_lik), np.array(ours_med), np.array(ours_pz)
    K, R = ours_lik.shape[0], ref_lik.shape[0]
    # marginal log-likelihood estimates per model: columns loglik_Null, loglik_Acc, loglik_Full of _elem_lik.txt
    for m in range(3):
        a, b = ours_lik[:, :, m], ref_lik[:, :, m]
        se = np.sqrt(a.var(axis=0, ddof=1) / K + b.var(axis=0, ddof=1) / R)
        assert np.all(np.abs(a.mean(axis=0) - b.mean(axis=0)) <= 4.0 * se + 0.75), (m, a.mean(axis=0), b.mean(axis=0), se)
    # Bayes factors and the accelerated call set (logBF1 > 5 and logBF2 > 5)
    bf1_o, bf2_o = ours_lik[:, :, 1] - ours_lik[:, :, 0], ours_lik[:, :, 1] - ours_lik[:, :, 2]
    bf1_r, bf2_r = ref_lik[:, :, 3], ref_lik[:, :, 4]
    for c in range(C):
        clear = abs(bf1_r[:, c].mean() - 5) > 3 and abs(bf2_r[:, c].mean() - 5) > 3
        if clear:
            call_r = bf1_r[:, c].mean() > 5 and bf2_r[:, c].mean() > 5
            call_o = bf1_o[:, c].mean() > 5 and bf2_o[:, c].mean() > 5
            assert call_r == call_o, (c, bf1_r[:, c].mean(), bf1_o[:, c].mean(), bf2_r[:, c].mean(), bf2_o[:, c].mean())
    assert (bf1_r.mean(axis=0) > 5).any()  # the fixture contains accelerated elements
    # rate medians (n, c, g, l, l2) and posterior P(Z = k) per node, averaged over seeds
    for m in range(3):
        med_r = ref_post[m][:, :, :5].mean(axis=0)
        med_o = ours_med[:, :, m, :].mean(axis=0)
        assert np.all(np.abs(med_o - med_r) <= 0.25 * np.abs(med_r) + 0.05), (m, med_o, med_r)
        pz_r = ref_post[m][:, :, 5:].reshape(R, C, N, 4).mean(axis=0)
        pz_o = ours_pz[:, :, m].mean(axis=0)
        assert np.max(np.abs(pz_o - pz_r)) < 0.2, (m, np.max(np.abs(pz_o - pz_r)))
        assert np.mean(np.abs(pz_o - pz_r)) < 0.02
    base.close()


def test_drop_in_binary_writes_reference_formats(native_lib):
    case = Case("ratite43", extra_params=dict(VERBOSE=1))
    assert os.path.exists(pab.BIN_PATH)
    # zero-argument probe of the driver (core.py:47-110): exit code 1, reference's message, no CUDA needed
    r0 = subprocess.run([pab.BIN_PATH], capture_output=True, text=True)
    assert r0.returncode == 1 and "Please specify the path of the parameter file." in r0.stdout
    r = subprocess.run([pab.BIN_PATH, case.cfg], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    log = r.stdout
    tline = [l for l in log.split("\n") if l.startswith("time used:")]
    assert tline and tline[0].split()[2].isdigit()  # phyloacc_post.py:176-179 reads the third token as an int
    out = os.path.join(case.dir, "out")
    ref = case.golden_refruns()[0]
    for fn in ("sim_elem_lik.txt", "sim_rate_postZ_M0.txt", "sim_rate_postZ_M1.txt", "sim_rate_postZ_M2.txt", "sim_M0_elem_Z.txt",
               "sim_M1_elem_Z.txt", "sim_M2_elem_Z.txt"):
        with open(os.path.join(out, fn)) as f:
            got = f.read()
        hg, rg = parse_table(got)
        hr, rr = parse_table(ref[fn])
        assert hg == hr, fn  # identical header line (column names are what phyloacc_post.py keys on)
        # the reference writes _rate_postZ rows in OpenMP completion order (bpp_c2.cpp:774-784); compare as sets of ids
        assert len(rg) == len(rr) and sorted(x[0] for x in rg) == sorted(x[0] for x in rr), fn
        assert all(len(a) == len(b) for a, b in zip(rg, rr)), fn
    for fn in ("sim_hyper.txt", "sim_species_names.txt"):
        assert os.path.exists(os.path.join(out, fn))
    with open(os.path.join(out, "sim_species_names.txt")) as f:
        assert f.read().split() == case.tree.names
    # VERBOSE traces: one file per (model, element) with the reference's header
    tr = os.path.join(out, "sim_mcmc_trace_M0_0.txt")
    assert os.path.exists(tr)
    with open(tr) as f:
        lines = f.read().split("\n")
    assert lines[0].startswith("loglik\trate_n\trate_c\tgrate\tlrate\tlrate2\t")
    assert len([l for l in lines[1:] if l]) == case.P.BURNIN + case.P.MCMC
    # logBF columns are consistent with the log-likelihood columns (bpp.cpp:578-581)
    with open(os.path.join(out, "sim_elem_lik.txt")) as f:
        h, rows = parse_table(f.read())
    for row in rows:
        v = [float(x) for x in row[2:]]
        if np.isfinite(v[:5]).all():
            assert abs((v[1] - v[0]) - v[3]) < 1e-4 * max(1, abs(v[3])) and abs((v[1] - v[2]) - v[4]) < 1e-4 * max(1, abs(v[4]))
    case.close()


def _run_binary(cfg, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([pab.BIN_PATH, cfg], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def _read_outputs(case):
    out = os.path.join(case.dir, "out")
    files = {}
    for fn in sorted(os.listdir(out)):
        with open(os.path.join(out, fn), "rb") as f:
            files[fn] = f.read()
    return files


@pytest.mark.parametrize("name,extra", [("ratite43", dict(VERBOSE=1)), ("tiny8", dict(SAMPLE_HYPER=1, CHAIN=3))])
def test_sharded_run_writes_byte_identical_files(native_lib, name, extra):
    """pa_run_from_config spreads the ids over all GPUs of the box (one engine + host thread per shard, pa_app.cpp).  Chains are keyed
    by the batch-wide element number, so every output file -- traces included, SAMPLE_HYPER's outer iterations included -- must be
    byte-identical whatever the number of shards.  With one GPU the shards share it (PA_DEVICE_LIST=0,0,0); with several they run on
    different GPUs."""
    import torch

    one = Case(name, extra_params=extra)
    _run_binary(one.cfg, dict(PA_DEVICES="1"))
    ref = _read_outputs(one)
    assert any(k.endswith("_elem_lik.txt") for k in ref)
    n_gpu = torch.cuda.device_count()
    lists = ["0,0,0"] + ([",".join(str(d) for d in range(min(n_gpu, 4)))] if n_gpu > 1 else [])
    for lst in lists:
        many = Case(name, extra_params=extra)
        log = _run_binary(many.cfg, dict(PA_DEVICE_LIST=lst, PA_MIN_ELEMS_PER_DEVICE="1"))
        assert "# element shards = " in log
        got = _read_outputs(many)
        assert sorted(got) == sorted(ref)
        for fn in ref:
            assert got[fn] == ref[fn], (lst, fn)
        many.close()
    one.close()
