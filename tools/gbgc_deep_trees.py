import os, sys, time, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np
import phyloacc_b200 as pab
from phyloacc_b200 import synth
for S in (110, 240):
    tree = synth.make_tree(S=S, seed=9, n_out=max(2, S // 6), n_target=max(2, S // 5))
    batch = synth.simulate(tree, C=6, length=200, seed=5, gap_frac=0.01, missing_frac=0.03, ambig_frac=0.002)
    d = tempfile.mkdtemp()
    cfg = synth.write_files(batch, d, prefix="sim", params=dict(BURNIN=5, MCMC=10, CHAIN=1, GAPCHAR="-"))
    eng = pab.Engine.from_config(cfg, device=0)
    try:
        t = time.time()
        out = eng.gbgc_run(list(range(6)), tree.targets, traces=True, num_burn=10, num_mcmc=20, seed=3)
        print("taxa", S, "ok: chain-iterations", out["chain_iterations"], "kernel ms %.1f" % out["kernel_ms"], "loglik finite:", np.isfinite(out["loglik_max"]).all())
        # every recorded state replays to its recorded likelihood
        el, Z, B, sn, sc, gb, want = [], [], [], [], [], [], []
        for i in range(6):
            if out["filtered"][i]: continue
            for mdl in range(3):
                for m in (0, 7, 29):
                    el.append(i); Z.append(np.maximum(out["trace_z"][i, mdl, m], 0)); B.append(out["trace_b"][i, mdl, m])
                    sc.append(out["trace"][i, mdl, 2, m]); sn.append(out["trace"][i, mdl, 3, m]); gb.append(out["trace"][i, mdl, 4, m]); want.append(out["trace"][i, mdl, 0, m])
        got = eng.gbgc_replay(el, np.asarray(Z), np.asarray(B), sn, sc, gb)
        print("   replay max rel err", np.abs(got / np.asarray(want) - 1).max())
    except Exception as e:
        print("taxa", S, "error:", e)
    eng.close()
