#!/usr/bin/env python
"""bench.py -- throughput of the PhyloAcc-ST hot path (per-element MCMC) on B200, BASELINE.json's metric:
element x MCMC-iterations / second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2|cfg3|cfg4] [--full-run]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs[0..2]; `config.workload` names the one that ran, cfg2 is the default / headline):
  cfg1  simu_500_200 shape: 43-taxon ratite-shaped tree, 500 elements x 200 bp (one step = the whole 500-element job)
  cfg2  ratite-genome scale: 43 taxa, element lengths lognormal(median 150, sigma 0.8) clipped to [50, 2000] bp, 2048 elements
        resident per GPU per step
  cfg3  240-taxon Zoonomia-shaped tree x 200 bp (deep-tree layout), 592 elements resident per GPU per step
  cfg4  PhyloAcc-ST-GBGC: 43 taxa x 200 bp simulated with a gBGC bias ~ Gamma(4, 0.8) on 20 % of the branches,
        2048 elements resident per GPU per step, M0 + M1 + M2 x (200 + 800) iterations (the variant's defaults) through
        pa_engine_gbgc_run; reference = the unmodified GBGC sources (oracle/_ref/PhyloAcc-GBGC-ref)
A "step" is one pass of the hot path over one resident batch: every element runs its M0, M1 and M2 chains for BURNIN + MCMC = 1500
iterations on the device.
value   = elements x 3 x 1500 / device time of the chain kernel, inputs resident in HBM (CUDA events on the launching stream);
e2e     = the same count / wall time of the public C-ABI call sequence with HOST buffers, over the same K steps:
          pa_engine_load_alignment (integer front-end + H2D) + pa_engine_run (chains + summaries + D2H of one record per chain),
          issued on two handles per GPU by two host threads so that the front-end of step k+1 overlaps the chains of step k;
roofline: FP64 FMA (north_star) -- algorithmic FMA per SURVEY.md §8d from the engine's work counters over the FP64 FMA peak of this
          GPU measured live (MEASURED_PEAKS.json has no FP64 entry; the theoretical 148 SM x 64 DFMA/clk is printed beside it);
cpu_baseline / --impl reference: the UNMODIFIED reference sources built against the compat shim (oracle/_ref/PhyloAcc-ST-ref) on all
          host cores, on a bounded STRATIFIED sample of the same workload: every step takes a fresh slice of >= 8 elements per core
          whose lengths are evenly spaced quantiles of the workload's length distribution, with short chains (per-iteration cost
          does not depend on chain length).  Two shapes are timed -- one process with NUM_THREAD = cores (OpenMP, schedule(guided))
          and the reference's native deployment shape, `cores` concurrent single-thread processes over length-balanced batches
           -- and the FASTER one is reported, named in `cpu_baseline.shape`.
--full-run: additionally run BASELINE's whole cfg2 job (280k elements, sharded over the ranks by the cost model of shard.py) once
          through the public API in resident waves and report it under "full_run".
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "element×MCMC-iter/sec (PhyloAcc-ST)"
METRIC_GBGC = "element×MCMC-iter/sec (PhyloAcc-ST-GBGC)"
UNIT = "element-iterations/s"
BURNIN, MCMC = 500, 1000  # driver defaults (params.py:130-131 of the reference)

WORKLOADS = {
    "cfg1": dict(name="cfg1: PhyloAcc-ST simu_500_200 shape, 43 taxa, 500 elements x 200 bp, M0+M1+M2 x (500+1000) iterations",
                 taxa=43, length=200, elems=500, ref_elems_per_core=8, ref_chain=(20, 40)),
    "cfg2": dict(name="cfg2: PhyloAcc-ST ratite-genome scale, 43 taxa, element lengths lognormal(150,0.8) in [50,2000] bp, M0+M1+M2 x (500+1000) iterations",
                 taxa=43, length="cfg2", elems=int(os.environ.get("PA_BENCH_ELEMS", "2048")), ref_elems_per_core=8, ref_chain=(20, 40)),
    "cfg3": dict(name="cfg3: PhyloAcc-ST on a 240-taxon Zoonomia-shaped tree, elements x 200 bp, M0+M1+M2 x (500+1000) iterations",
                 taxa=240, length=200, elems=592, ref_elems_per_core=4, ref_chain=(10, 20)),
    "cfg4": dict(name="cfg4: PhyloAcc-ST-GBGC, 43 taxa, elements x 200 bp, gBGC bias ~ Gamma(4,0.8) on 20% of the branches, M0+M1+M2 x (200+800) iterations",
                 taxa=43, length=200, elems=int(os.environ.get("PA_BENCH_ELEMS", "2048")), ref_elems_per_core=3, ref_chain=(20, 40), gbgc=True),
}
GBGC_BURNIN, GBGC_MCMC = 200, 800  # defaults of the GBGC binary (GBGC/SRC/main.cpp:52-53)
FULL_RUN_ELEMS = 280000  # BASELINE.json configs[1]


def make_workload_tree(wl):
    from phyloacc_b200 import synth

    if wl["taxa"] == 43:
        return synth.make_tree()
    S = wl["taxa"]
    return synth.make_tree(S=S, seed=9, n_out=max(2, S // 6), n_target=max(2, S // 5))


def _clock_sampler(stop, out, dev):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(dev), f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def _clock_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    reasons = set()
    for s in samples:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None,
            "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ the CPU reference
def stratified_lengths(wl, n, slice_index, n_slices):
    """Lengths of slice `slice_index` of a stratified sample of n * n_slices elements: evenly spaced quantiles of the workload's
    length distribution, interleaved so that every slice is itself stratified and no two slices share an element."""
    if wl["length"] != "cfg2":
        return np.full(n, int(wl["length"]), dtype=np.int32)
    from statistics import NormalDist

    tot = n * n_slices
    q = (np.arange(n) * n_slices + (slice_index % n_slices) + 0.5) / tot
    z = np.array([NormalDist().inv_cdf(float(x)) for x in q])
    return np.clip(np.round(150.0 * np.exp(0.8 * z)), 50, 2000).astype(np.int32)


def _lpt_batches(lengths, k):
    """Longest-processing-time split of element indices into k batches of near-equal total length (the steady state of the
    reference's deployment -- many 50-locus batches streaming through `cores` workers -- has no tail; a bounded sample must not
    invent one)."""
    order = np.argsort(-np.asarray(lengths), kind="stable")
    loads = np.zeros(k)
    out = [[] for _ in range(k)]
    for i in order:
        j = int(np.argmin(loads))
        out[j].append(int(i))
        loads[j] += lengths[i]
    return [sorted(b) for b in out if b]


def _count_done(path):
    try:
        with open(path) as f:
            return max(0, len(f.read().strip().split("\n")) - 1)
    except OSError:
        return 0


def reference_step(wl, tree, cores, slice_index, n_slices, shape):
    """One bounded sample of the workload through the unmodified reference binary (shim build).  shape = "omp" (one process,
    NUM_THREAD = cores) or "procs" (cores concurrent single-thread processes, length-balanced batches).
    Returns (element-iterations/s, wall seconds, elements done, description)."""
    from phyloacc_b200 import synth

    ref_bin = os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-ST-ref")
    burn, mc = wl["ref_chain"]
    n_el = max(32, wl["ref_elems_per_core"] * cores)
    lengths = stratified_lengths(wl, n_el, slice_index, n_slices)
    batch = synth.simulate(tree, lengths=lengths, seed=4242 + slice_index)
    tmp = tempfile.mkdtemp(prefix="pa_bench_ref_")
    try:
        if not os.path.exists(ref_bin):
            return _port_step(wl, tree, batch, burn, mc, tmp)
        prm = dict(BURNIN=burn, MCMC=mc, CHAIN=1, ADAPT_FREQ=500, SEED=1 + slice_index)
        if shape == "omp":
            cfg = synth.write_files(batch, tmp, prefix="ref", params=dict(prm, NUM_THREAD=cores))
            t0 = time.time()
            with open(os.path.join(tmp, "log.txt"), "w") as lf:
                subprocess.run([ref_bin, cfg], stdout=lf, stderr=subprocess.STDOUT, check=True)
            dt = time.time() - t0
            n_done = _count_done(os.path.join(tmp, "out", "ref_rate_postZ_M0.txt"))
        else:
            batches = _lpt_batches(lengths, cores)
            cfgs = []
            for b, ids in enumerate(batches):
                d = os.path.join(tmp, f"b{b}")
                cfgs.append(synth.write_files(batch, d, prefix="ref", ids=ids, params=dict(prm, NUM_THREAD=1)))
            t0 = time.time()
            procs = []
            for b, cfg in enumerate(cfgs):
                lf = open(os.path.join(tmp, f"b{b}", "log.txt"), "w")
                procs.append((subprocess.Popen([ref_bin, cfg], stdout=lf, stderr=subprocess.STDOUT), lf))
            for p, lf in procs:
                p.wait()
                lf.close()
            dt = time.time() - t0
            n_done = sum(_count_done(os.path.join(tmp, f"b{b}", "out", "ref_rate_postZ_M0.txt")) for b in range(len(cfgs)))
        value = n_done * 3 * (burn + mc) / dt
        what = (f"{n_done} {wl['name'].split(':')[0]} elements (stratified slice {slice_index} of {n_slices}, lengths {int(lengths.min())}-{int(lengths.max())} bp, "
                f"mean {lengths.mean():.0f}) x 3 models x ({burn}+{mc}) iterations, "
                + (f"one process NUM_THREAD={cores}" if shape == "omp" else f"{len(batches)} concurrent single-thread processes, length-balanced batches")
                + f", {dt:.1f} s wall (reference sources + Armadillo/GSL compat shim, g++ -O2 -fopenmp)")
        return value, dt, n_done, what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def gbgc_reference_step(wl, tree, cores, slice_index):
    """One bounded sample of cfg4 through the unmodified PhyloAcc-ST-GBGC binary (shim build): `cores` concurrent single-thread
    processes, each with its own ids of one shared input.  Returns (element-iterations/s, seconds, elements done, description)."""
    from phyloacc_b200 import synth

    ref_bin = os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-GBGC-ref")
    burn, mc = wl["ref_chain"]
    n_el = max(8, wl["ref_elems_per_core"] * cores)
    batch = synth.simulate(tree, C=n_el, length=wl["length"], seed=5150 + slice_index, gbgc={})
    tmp = tempfile.mkdtemp(prefix="pa_bench_gbgc_")
    try:
        if not os.path.exists(ref_bin):
            return _gbgc_port_step(wl, tree, batch, burn, mc)
        prm = dict(BURNIN=burn, MCMC=mc, CHAIN=1, ADAPT_FREQ=100, SEED=1 + slice_index, GAPCHAR="-", VERBOSE=0, NUM_THREAD=1)
        cfgs = []
        for b in range(cores):
            ids = list(range(b, n_el, cores))
            if ids:
                cfgs.append(synth.write_files(batch, os.path.join(tmp, f"b{b}"), prefix="ref", ids=ids, params=prm))
        t0 = time.time()
        procs = []
        for b, cfg in enumerate(cfgs):
            lf = open(os.path.join(os.path.dirname(cfg), "log.txt"), "w")
            procs.append((subprocess.Popen([ref_bin, cfg], stdout=lf, stderr=subprocess.STDOUT), lf))
        for p, lf in procs:
            p.wait()
            lf.close()
        dt = time.time() - t0
        n_done = sum(_count_done(os.path.join(os.path.dirname(cfg), "out", "ref_rate_postZ_1.txt")) for cfg in cfgs)
        what = (f"{n_done} cfg4 elements (43 taxa x {wl['length']} bp, simulated with gBGC) x 3 models x ({burn}+{mc}) iterations, {len(cfgs)} concurrent "
                f"single-thread processes, {dt:.1f} s wall (unmodified PhyloAcc-ST-GBGC sources + Armadillo/GSL compat shim, g++ -O2)")
        return n_done * 3 * (burn + mc) / dt, dt, n_done, what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _gbgc_port_step(wl, tree, batch, burn, mc):
    """No GBGC reference binary on this box: time the numpy restatement of the chain on one core (cpu_baseline.kind = "port")."""
    from oracle import frontend as fe, gbgc  # cpu_baseline leg only
    from phyloacc_b200 import synth

    tmp = tempfile.mkdtemp(prefix="pa_bench_gbgc_port_")
    try:
        cfg = synth.write_files(batch, tmp, prefix="ref", params=dict(BURNIN=burn, MCMC=mc, CHAIN=1, GAPCHAR="-"))
        P = fe.load_params(cfg)
        otree = fe.parse_mod(P.PHYTREE_FILE)
        names, X = fe.load_fasta(P.ALIGN_FILE)
        _, pos = fe.load_bed(P.SEG_FILE)
        starts, lens = fe.element_layout(pos)
        model = fe.build_model(otree, P)
        X = fe.reorder_alignment(otree, names, X)
        par = gbgc.GbgcParams(num_burn=burn, num_mcmc=mc)
        t0 = time.time()
        n_done = 0
        for c in range(min(batch.C, 2)):
            el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
            if el.filter:
                continue
            for resZ in (0, 2, 1):
                gbgc.GbgcChain(np.asarray(otree.Q).reshape(4, 4), otree.pi, el, otree.N, otree.S, resZ, tree.targets, model.upper, par, 1, c * 4 + resZ).run()
            n_done += 1
        dt = time.time() - t0
        return n_done * 3 * (burn + mc) / dt, dt, n_done, f"numpy port of the GBGC chain, {n_done} elements x 3 models x ({burn}+{mc}) iterations on 1 core, {dt:.1f} s"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def gbgc_cpu_baseline(wl, tree, cores, slices=(1, 2)):
    have_bin = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-GBGC-ref"))
    units, secs, what = 0.0, 0.0, ""
    for sl in slices:
        v, dt, _, what = gbgc_reference_step(wl, tree, cores, sl)
        units += v * dt
        secs += dt
    return dict(value=units / secs, unit=UNIT, cores=cores if have_bin else 1, kind="reference" if have_bin else "port", shape="procs",
                sample=f"{len(slices)} samples; last: {what}")


def run_gbgc_reference_arm(args, wl, rank):
    if rank != 0:
        return
    tree = make_workload_tree(wl)
    cores = os.cpu_count() or 1
    have_bin = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-GBGC-ref"))
    for k in range(args.warmup):
        gbgc_reference_step(wl, tree, cores, k)
    units, secs, what = 0.0, 0.0, ""
    for k in range(args.steps):
        v, dt, _, what = gbgc_reference_step(wl, tree, cores, args.warmup + k)
        units += v * dt
        secs += dt
    v = units / secs
    base = dict(value=v, unit=UNIT, cores=cores if have_bin else 1, kind="reference" if have_bin else "port", shape="procs",
                sample=f"{args.steps} distinct samples ({args.steps + args.warmup} incl. warm-up); last: {what}")
    line = {"impl": "reference", "metric": METRIC_GBGC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": wl["name"], "sample": base["sample"]}, "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main_gbgc(args, wl, rank, local_rank, world):
    """cfg4: the PhyloAcc-ST-GBGC chain (pa_engine_gbgc_run).  Same contract as the ST workloads: `value` from the chain kernel's CUDA
    events (inputs resident), `e2e` = pa_engine_load_alignment + pa_engine_gbgc_run with host buffers, wall clock."""
    import torch
    import torch.distributed as dist

    import phyloacc_b200 as pab
    from phyloacc_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tree = make_workload_tree(wl)
    batch = synth.simulate(tree, C=wl["elems"], length=wl["length"], seed=4000 + rank, gbgc={})
    prm = pab.default_params()
    prm.num_thread = max(1, (os.cpu_count() or 1) // max(world, 1))
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup, conserve=tree.conserve,
                            device=local_rank)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    Xn = torch.from_numpy(batch.X).pin_memory().numpy()
    eng.load_alignment(Xn, batch.elem_start, batch.elem_len)
    ids = np.arange(batch.C, dtype=np.int32)
    T = GBGC_BURNIN + GBGC_MCMC
    kw = dict(num_burn=GBGC_BURNIN, num_mcmc=GBGC_MCMC, seed=1 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks, stop = [], threading.Event()
    out = None
    for _ in range(args.warmup):
        out = eng.gbgc_run(ids, tree.targets, **kw)
    n_active = int((out["filtered"] == 0).sum())
    units_per_step = n_active * 3 * T
    sampler = threading.Thread(target=_clock_sampler, args=(stop, clocks, local_rank), daemon=True)
    sampler.start()
    barrier()
    kern_ms = []
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k)
        torch.cuda.synchronize()
        out = eng.gbgc_run(ids, tree.targets, **kw)  # kernel_ms: CUDA events around the chain kernel on the engine's stream
        kern_ms.append(out["kernel_ms"])
    barrier()
    ms_step = float(np.mean(kern_ms))
    barrier()
    t0 = time.time()
    for k in range(args.steps):
        eng.load_alignment(Xn, batch.elem_start, batch.elem_len)
        out = eng.gbgc_run(ids, tree.targets, **kw)
    barrier()
    e2e_ms = (time.time() - t0) * 1e3 / args.steps
    stop.set()
    sampler.join(timeout=2)
    t_ms = torch.tensor([ms_step, e2e_ms], device="cuda", dtype=torch.float64)
    units = torch.tensor([float(units_per_step)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        summ = torch.from_numpy(np.ascontiguousarray(out["loglik"])).cuda()
        gathered = [torch.empty_like(summ) for _ in range(world)]
        dist.all_gather(gathered, summ)
    value = float(units[0]) / (float(t_ms[0]) * 1e-3)
    e2e_value = float(units[0]) / (float(t_ms[1]) * 1e-3)
    if rank == 0:
        # algorithmic FMA of one step: 36 per (internal node, pattern) of an upward pass, 48 per (branch, pattern) of a downward pass
        # (16 message + 16 joint + 16 weighted accumulation), 96 per (branch, class) matrix (80 for P(t) as in §8d + its 16-term emission)
        fma = 36.0 * out["node_pattern_updates"] + 48.0 * out["branch_pattern_updates"] + 96.0 * out["pt_builds"]
        achieved = 2.0 * fma / (ms_step * 1e-3) / 1e12
        peak = pab.fp64_peak_tflops(local_rank, 400)
        its = max(1, out["chain_iterations"])
        names = "setup monitor downward init_emissions init_messages pt_builds emissions messages zb_draws upward elbo_mh commit".split()
        sm_clock = (_clock_summary(clocks)["sm_mhz"] or 1965) * 1e6
        lr = np.asarray(out["loglik"])
        line = {"metric": METRIC_GBGC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(t_ms[0]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["name"], "elements_per_gpu_per_step": wl["elems"], "active_elements_rank0": n_active, "iterations_per_chain": T,
                           "models": 3, "l2": "flushed between timed steps (256 MiB write)",
                           "parallelism": f"elements sharded over {world} GPU(s), no data-path collective, one all_gather of summaries"},
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": None,
                             "peak_source": "measured live: register-resident DFMA micro-benchmark (pa_device_fp64_peak)",
                             "peak_theoretical": 148 * 64 * 2 * sm_clock / 1e12, "algorithmic_fma_per_step": fma, "kernel": "gbgc_chain_kernel",
                             "phase_kcycles_per_chain_iteration": {n: round(c / its / 1e3, 1) for n, c in zip(names, out["phase_cycles"])}},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(batch.X.nbytes), "d2h_bytes_per_step": int(out["d2h_bytes"]), "steps": args.steps},
                "gpu_launches": args.steps * int(out["kernel_launches"]), "clocks": _clock_summary(clocks),
                "work": {k: int(out[k]) for k in ("node_pattern_updates", "branch_pattern_updates", "pt_builds", "chain_iterations")},
                "accelerated_calls_rank0": int(np.sum((lr[:, 1] - lr[:, 0] > 5) & (lr[:, 1] - lr[:, 2] > 5) & (out["filtered"] == 0)))}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = gbgc_cpu_baseline(wl, tree, os.cpu_count() or 1)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _port_step(wl, tree, batch, burn, mc, tmp):
    """No reference binary on this box: time the oracle port on one core (cpu_baseline.kind = "port")."""
    from oracle import coracle, frontend as fe  # cpu_baseline leg only
    from phyloacc_b200 import synth

    cfg = synth.write_files(batch, tmp, prefix="ref", params=dict(BURNIN=burn, MCMC=mc, CHAIN=1, ADAPT_FREQ=500, SEED=1))
    P = fe.load_params(cfg)
    otree = fe.parse_mod(P.PHYTREE_FILE)
    names, X = fe.load_fasta(P.ALIGN_FILE)
    _, pos = fe.load_bed(P.SEG_FILE)
    starts, lens = fe.element_layout(pos)
    model = fe.build_model(otree, P)
    X = fe.reorder_alignment(otree, names, X)
    t0 = time.time()
    n_done = 0
    for c in range(min(batch.C, 4)):
        el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
        if el.filter:
            continue
        h = coracle.ElementHandle(model, el)
        for resZ in (0, 2, 1):
            h.chain(resZ, 1, c * 4 + resZ)
        n_done += 1
    dt = time.time() - t0
    return n_done * 3 * (burn + mc) / dt, dt, n_done, f"oracle port, {n_done} elements x 3 models x ({burn}+{mc}) iterations on 1 core, {dt:.1f} s"


def reference_calibrate(wl, tree, cores, n_slices):
    """Both shapes on slice 0; returns (best shape, {shape: rate})."""
    rates = {}
    have_bin = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-ST-ref"))
    for shape in (("omp", "procs") if have_bin else ("omp",)):
        rates[shape] = reference_step(wl, tree, cores, 0, n_slices, shape)[0]
    best = max(rates, key=rates.get)
    return best, rates


def cpu_baseline_block(wl, tree, cores, n_slices=25, slices=(1, 2)):
    """The cpu_baseline object of our arm's line: calibrate the shape on slice 0, then time `slices` with the faster one."""
    have_bin = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-ST-ref"))
    best, rates = reference_calibrate(wl, tree, cores, n_slices)
    tot_units, tot_dt, what = 0.0, 0.0, ""
    for s in slices:
        v, dt, n_done, what = reference_step(wl, tree, cores, s, n_slices, best)
        tot_units += v * dt
        tot_dt += dt
    return dict(value=tot_units / tot_dt, unit=UNIT, cores=cores if have_bin else 1, kind="reference" if have_bin else "port", shape=best,
                shape_rates={k: round(v, 1) for k, v in rates.items()},
                sample=f"{len(slices)} slices; last: {what}")


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    tree = make_workload_tree(wl)
    cores = os.cpu_count() or 1
    have_bin = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-ST-ref"))
    n_slices = max(2, args.steps + args.warmup)
    # warm-up steps alternate the two shapes on their own slices; the timed steps use the faster one (with fewer than two warm-up
    # steps both shapes are calibrated on slice 0 first, untimed)
    rates = {}
    shapes = ("omp", "procs") if have_bin else ("omp",)
    if args.warmup < len(shapes):
        best, rates = reference_calibrate(wl, tree, cores, n_slices)
    for k in range(args.warmup):
        sh = shapes[k % len(shapes)]
        v = reference_step(wl, tree, cores, k, n_slices, sh)[0]
        rates[sh] = max(rates.get(sh, 0.0), v)
    best = max(rates, key=rates.get)
    units, secs, what = 0.0, 0.0, ""
    for k in range(args.steps):
        v, dt, n_done, what = reference_step(wl, tree, cores, args.warmup + k, n_slices, best)
        units += v * dt
        secs += dt
    v = units / secs
    base = dict(value=v, unit=UNIT, cores=cores if have_bin else 1, kind="reference" if have_bin else "port", shape=best,
                shape_rates={k: round(x, 1) for k, x in rates.items()},
                sample=f"{args.steps} distinct stratified slices ({args.steps + args.warmup} incl. warm-up); last: {what}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": wl["name"], "sample": base["sample"]}, "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def newest_traffic(workload_key):
    """DRAM bytes per chain-iteration of the chain kernel from the newest committed ncu capture of this workload
    (profiles/r<round>_v<k>_traffic[_cfgN].json); (None, None) when there is none."""
    import glob
    import re

    suffix = "" if workload_key == "cfg2" else f"_{workload_key}"
    cands = [q for q in glob.glob(os.path.join(ROOT, "profiles", f"r*_v*_traffic{suffix}.json"))]
    if not cands:
        return None, None
    cands.sort(key=lambda q: [int(x) for x in re.findall(r"\d+", os.path.basename(q))])
    try:
        with open(cands[-1]) as f:
            return float(json.load(f)["dram_bytes_per_chain_iteration"]), os.path.basename(cands[-1])
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--elems", type=int, default=0, help="elements resident per GPU per step (default: the workload's)")
    ap.add_argument("--full-run", action="store_true", help="also run the whole 280k-element cfg2 job once through the public API")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.elems:
        wl["elems"] = args.elems
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if wl.get("gbgc"):
        if args.impl == "reference":
            run_gbgc_reference_arm(args, wl, rank)
        else:
            main_gbgc(args, wl, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist

    import phyloacc_b200 as pab
    from phyloacc_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic shard of this rank (elements are independent: contiguous element ranges per GPU, no data-path collective)
    tree = make_workload_tree(wl)
    batch = synth.simulate(tree, C=wl["elems"], length=wl["length"], seed=1000 + rank)
    prm = pab.default_params()
    prm.burnin, prm.mcmc, prm.adapt_freq, prm.seed = BURNIN, MCMC, 500, 1 + rank
    prm.num_thread = max(1, (os.cpu_count() or 1) // max(world, 1))
    eng = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                            conserve=tree.conserve, device=local_rank)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    X_host = torch.from_numpy(batch.X).pin_memory()
    Xn = X_host.numpy()
    eng.load_alignment(Xn, batch.elem_start, batch.elem_len)
    ids = np.arange(batch.C, dtype=np.int32)
    n_active = int(sum(not eng.element_info(int(c))["filtered"] for c in ids))
    T = BURNIN + MCMC
    units_per_step = n_active * 3 * T
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # L2 flush between timed steps (126 MB L2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only throughput, inputs resident in HBM
    clocks, stop = [], threading.Event()
    for _ in range(args.warmup):
        eng.sample(ids)
        torch.cuda.synchronize()
    sampler = threading.Thread(target=_clock_sampler, args=(stop, clocks, local_rank), daemon=True)
    sampler.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k)
            ev[k][0].record(stream)
            eng.sample(ids)
            ev[k][1].record(stream)
    barrier()
    wall_ms = (time.time() - t_wall0) * 1e3
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    ms_step = float(np.mean(kern_ms))
    cnt = eng.collect() and eng.counters()

    # ---- end to end through the C ABI with host buffers, over the same number of steps (front-end + H2D + chains + summaries + D2H).
    # Two handles per GPU, used alternately by two host threads: the integer front-end and upload of step k+1 (pa_engine_load_alignment
    # on the second handle) overlap the chains of step k, as a caller with a queue of batches would arrange it.  Every step still makes
    # both calls with host buffers; all of it is inside the timed region; the two handles' kernels share the GPU.
    eng_b = pab.Engine.create(tree.parent, tree.children, tree.dist, tree.pi, tree.Q, params=prm, outgroup=tree.outgroup,
                              conserve=tree.conserve, device=local_rank)
    stream_b = torch.cuda.Stream()
    eng_b.set_stream(stream_b.cuda_stream)
    engines = [eng, eng_b]
    for e in engines:  # untimed: buffers of both handles allocated once
        e.load_alignment(Xn, batch.elem_start, batch.elem_len)
        e.run(ids)
    last_res = [None, None]

    def e2e_worker(slot):
        torch.cuda.set_device(local_rank)
        for k in range(slot, args.steps, 2):
            engines[slot].load_alignment(Xn, batch.elem_start, batch.elem_len)
            last_res[slot] = engines[slot].run(ids)

    barrier()
    t0 = time.time()
    workers = [threading.Thread(target=e2e_worker, args=(s,)) for s in range(2)]
    for w in workers:
        w.start()
    for w in workers:
        w.join()
    barrier()
    e2e_ms = (time.time() - t0) * 1e3 / args.steps
    res = last_res[0] if last_res[0] is not None else last_res[1]
    stop.set()
    sampler.join(timeout=2)
    h2d = int(batch.X.nbytes)
    d2h = int(eng.d2h_bytes()) if hasattr(eng, "d2h_bytes") else int(n_active * 3 * T * (6 * 8 + ((2 * tree.S - 1 + 15) // 16) * 4))
    n_acc = int(np.sum((res.logBF1 > 5) & (res.logBF2 > 5) & (res.filtered == 0)))

    # ---- optional: BASELINE's whole cfg2 job (280k elements over all ranks) through the public API in resident waves
    full = None
    if args.full_run:
        from phyloacc_b200 import shard

        wl2 = WORKLOADS["cfg2"]
        tree2 = tree if args.workload != "cfg3" else make_workload_tree(wl2)
        rng = np.random.default_rng(280)
        all_len = np.clip(np.round(150.0 * np.exp(0.8 * rng.standard_normal(FULL_RUN_ELEMS))), 50, 2000).astype(np.int32)
        mine = shard.shard_ranges(all_len.astype(np.float64) * (tree2.S - 1), world)[rank]
        my_len = all_len[mine.start:mine.stop]
        big = synth.simulate(tree2, lengths=my_len, seed=2800 + rank)
        eng2 = eng if args.workload != "cfg3" else pab.Engine.create(tree2.parent, tree2.children, tree2.dist, tree2.pi, tree2.Q, params=prm,
                                                                      outgroup=tree2.outgroup, conserve=tree2.conserve, device=local_rank)
        wave = 8192
        engs2 = [eng2, eng_b] if args.workload != "cfg3" else [eng2]
        totals = [[0, 0] for _ in engs2]
        waves = list(range(0, big.C, wave))

        def full_worker(slot):
            torch.cuda.set_device(local_rank)
            for w0 in waves[slot::len(engs2)]:
                w1 = min(big.C, w0 + wave)
                g0, g1 = int(big.elem_start[w0]), int(big.elem_start[w1 - 1] + big.elem_len[w1 - 1])
                Xw = np.ascontiguousarray(big.X[:, g0:g1])
                engs2[slot].load_alignment(Xw, big.elem_start[w0:w1] - g0, big.elem_len[w0:w1])
                r2 = engs2[slot].run(np.arange(w1 - w0, dtype=np.int32))
                totals[slot][0] += int((r2.filtered == 0).sum())
                totals[slot][1] += int(np.sum((r2.logBF1 > 5) & (r2.logBF2 > 5) & (r2.filtered == 0)))

        barrier()
        t0 = time.time()
        fw = [threading.Thread(target=full_worker, args=(s,)) for s in range(len(engs2))]
        for w in fw:
            w.start()
        for w in fw:
            w.join()
        n_full_active, n_full_acc = sum(t[0] for t in totals), sum(t[1] for t in totals)
        barrier()
        full_s = time.time() - t0
        tf = torch.tensor([full_s], device="cuda", dtype=torch.float64)
        nf = torch.tensor([float(n_full_active), float(n_full_acc), float(big.C)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            dist.all_reduce(nf, op=dist.ReduceOp.SUM)
        full = {"elements": int(nf[2]), "active_elements": int(nf[0]), "accelerated_calls": int(nf[1]), "seconds": float(tf[0]),
                "value": float(nf[0]) * 3 * T / float(tf[0]), "unit": UNIT, "waves_of": wave,
                "what": "whole cfg2 job through pa_engine_load_alignment + pa_engine_run with host buffers, elements sharded over the ranks by cost"}

    # ---- reduce over ranks: max time, summed units; one gather of per-element summaries (NCCL)
    t_ms = torch.tensor([ms_step, e2e_ms, wall_ms], device="cuda", dtype=torch.float64)
    units = torch.tensor([float(units_per_step)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        summ = torch.from_numpy(np.stack([res.logBF1, res.logBF2], axis=1)).cuda()
        gathered = [torch.empty_like(summ) for _ in range(world)]
        dist.all_gather(gathered, summ)  # the single gather of per-element summaries at the end of the run
    ms_step_all, e2e_ms_all = float(t_ms[0]), float(t_ms[1])
    total_units = float(units[0])
    value = total_units / (ms_step_all * 1e-3)
    e2e_value = total_units / (e2e_ms_all * 1e-3)

    if rank == 0:
        # roofline: algorithmic FMA = 36/node-pattern update + 4/root evaluation + 80/P(t) build + 4/history draw
        fma = 36.0 * cnt["node_pattern_updates"] + 4.0 * cnt["root_evals"] + 80.0 * cnt["pt_builds"] + 4.0 * cnt["history_draws"]
        achieved = 2.0 * fma / (ms_step * 1e-3) / 1e12
        peak = pab.fp64_peak_tflops(local_rank, 400)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_bytes = float(eng_blob_bytes(eng, ids)) + d2h  # alignment streamed once + result write-back
        # DRAM traffic of the chain kernel per launch: measured once under ncu on this workload (newest committed capture, bytes per
        # chain-iteration) and scaled to this launch's chain-iterations; null when no capture of this workload is committed
        per_iter, traffic_file = newest_traffic(args.workload)
        traffic = per_iter * float(cnt["chain_iterations"]) if per_iter else None
        sm_clock = (_clock_summary(clocks)["sm_mhz"] or 1965) * 1e6
        roof = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic,
                "traffic_source": (f"ncu dram__bytes_read+write of one bench-size chain_kernel launch (profiles/{traffic_file}), scaled by chain-iterations"
                                   if traffic_file else None),
                "peak_source": "measured live: register-resident DFMA micro-benchmark (pa_device_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "peak_theoretical": 148 * 64 * 2 * sm_clock / 1e12,
                "algorithmic_fma_per_step": fma, "kernel": "chain_kernel",
                "hbm": {"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                        "note": "algorithmic bytes (alignment read once + results written); the kernel's measured DRAM traffic is below"},
                # measured DRAM traffic of the launch (arena spills out of L2) over this run's kernel time: how far the kernel is from
                # being bandwidth-bound
                "dram_measured": ({"achieved": traffic / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": traffic / (ms_step * 1e-3) / 1e9 / hbm_peak} if traffic else None)}
        # BASELINE's second metric: full-tree-equivalent site likelihoods per second on UNCOMPRESSED columns
        # (sum over upward passes of columns x recomputed internal nodes / (S - 1)), rank 0's share x ranks
        site_liks = float(cnt["node_site_updates"]) / (tree.S - 1) / (ms_step * 1e-3) * world
        compression = float(cnt["node_site_updates"]) / max(1.0, float(cnt["node_pattern_updates"]))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step_all, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": wl["name"], "elements_per_gpu_per_step": wl["elems"], "active_elements_rank0": n_active,
                           "iterations_per_chain": T, "models": 3, "l2": "flushed between timed steps (256 MiB write)",
                           "parallelism": f"elements sharded over {world} GPU(s), no data-path collective, one all_gather of summaries"},
                "roofline": roof,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.steps,
                        "pipeline": "two handles per GPU on two host threads: front-end + upload of step k+1 overlap the chains of step k"},
                "pruned_site_liks": {"value": site_liks, "unit": "full-tree-equivalent site likelihoods/s", "columns_per_pattern": compression},
                "gpu_launches": int(cnt["kernel_launches"]) * args.steps,
                "clocks": _clock_summary(clocks),
                "work": {k: int(v) for k, v in cnt.items()}, "accelerated_calls_rank0": n_acc}
        if full is not None:
            line["full_run"] = full
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(wl, tree, os.cpu_count() or 1)
        print(json.dumps(line))
    eng_b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def eng_blob_bytes(eng, ids):
    # packed alignment bytes read by the chains: 4-bit codes of every pattern for every species
    tot = 0
    for c in ids[: min(len(ids), 256)]:
        info = eng.element_info(int(c))
        if not info["filtered"]:
            tot += info["P"] * eng.S // 2
    return tot * (len(ids) / max(1, min(len(ids), 256)))


if __name__ == "__main__":
    main()
