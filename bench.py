#!/usr/bin/env python
"""bench.py -- throughput of the PhyloAcc-ST hot path (per-element MCMC) on B200, BASELINE.json's metric:
element x MCMC-iterations / second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one resident batch of synthetic elements drawn from the cfg2 workload
(BASELINE.json configs[1]: 43-taxon ratite-shaped tree, element lengths lognormal(median 150, sigma 0.8) clipped to
[50, 2000] bp): every element runs its M0, M1 and M2 chains for BURNIN+MCMC = 1500 iterations on the device.
value   = elements x 3 x 1500 / device time of the chain kernel, inputs resident in HBM (CUDA events);
e2e     = the same count / time of the public C-ABI call sequence with HOST buffers: pa_engine_load_alignment
          (integer front-end + H2D) + pa_engine_run (chains + D2H of traces + Eval2 / summaries on the host, finished chains
          summarised while later ones are still being sampled);
roofline: FP64 FMA (north_star) -- algorithmic FMA per SURVEY.md §8d from the engine's work counters over the measured
          FP64 FMA peak of this GPU (micro-benchmark, MEASURED_PEAKS.json has no FP64 entry);
cpu_baseline / --impl reference: the UNMODIFIED reference sources built against the compat shim
          (oracle/_ref/PhyloAcc-ST-ref) on the host cores, on a bounded sample of the same workload.
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
UNIT = "element-iterations/s"
BURNIN, MCMC = 500, 1000  # driver defaults (params.py:130-131 of the reference)
ELEMS_PER_GPU = int(os.environ.get("PA_BENCH_ELEMS", "2048"))
WORKLOAD = "cfg2: PhyloAcc-ST ratite-genome scale, 43 taxa, element lengths lognormal(150,0.8) in [50,2000] bp, M0+M1+M2 x (500+1000) iterations"


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


def make_batch(n_elems, seed):
    from phyloacc_b200 import synth

    tree = synth.make_tree()
    batch = synth.simulate(tree, C=n_elems, length="cfg2", seed=seed)
    return tree, batch


def reference_throughput(tree, cores, seconds_target=20.0, label="reference"):
    """Time the unmodified reference (shim build) with NUM_THREAD = all host cores on a bounded sample of the workload:
    per-iteration cost does not depend on chain length, so the sample uses shorter chains (BURNIN 30 + MCMC 60)."""
    from phyloacc_b200 import synth

    ref_bin = os.path.join(ROOT, "oracle", "_ref", "PhyloAcc-ST-ref")
    burn, mc = 30, 60
    kind = "reference"
    if not os.path.exists(ref_bin):
        kind = "port"
    n_el = max(cores, 8) * 2
    batch = synth.simulate(tree, C=n_el, length="cfg2", seed=4242)
    tmp = tempfile.mkdtemp(prefix="pa_bench_ref_")
    try:
        if kind == "reference":
            cfg = synth.write_files(batch, tmp, prefix="ref", params=dict(BURNIN=burn, MCMC=mc, CHAIN=1, ADAPT_FREQ=500, SEED=1, NUM_THREAD=cores))
            t0 = time.time()
            with open(os.path.join(tmp, "log.txt"), "w") as lf:
                subprocess.run([ref_bin, cfg], stdout=lf, stderr=subprocess.STDOUT, check=True)
            dt = time.time() - t0
            n_done = 0
            with open(os.path.join(tmp, "out", "ref_rate_postZ_M0.txt")) as f:
                n_done = max(0, len(f.read().strip().split("\n")) - 1)
        else:
            from oracle import coracle, frontend as fe  # cpu_baseline leg only

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
            for c in range(min(n_el, 4)):
                el = fe.element_frontend(model, X, int(starts[c]), int(lens[c]))
                if el.filter:
                    continue
                h = coracle.ElementHandle(model, el)
                for resZ in (0, 2, 1):
                    h.chain(resZ, 1, c * 4 + resZ)
                n_done += 1
            dt = time.time() - t0
            cores = 1
        value = n_done * 3 * (burn + mc) / dt
        return dict(value=value, unit=UNIT, cores=cores, kind=kind,
                    sample=f"{n_done} cfg2 elements x 3 models x ({burn}+{mc}) iterations, NUM_THREAD={cores}, {dt:.1f} s wall "
                           f"(reference sources + Armadillo/GSL compat shim, g++ -O2 -fopenmp)"), dt
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from phyloacc_b200 import synth

    tree = synth.make_tree()
    cores = os.cpu_count() or 1
    vals, ms = [], []
    last = None
    for k in range(args.warmup + args.steps):
        base, dt = reference_throughput(tree, cores)
        if k >= args.warmup:
            vals.append(base["value"])
            ms.append(dt * 1e3)
        last = base
    v = float(np.mean(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": last["sample"]}, "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=ELEMS_PER_GPU, help="elements resident per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import phyloacc_b200 as pab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic shard of this rank (elements are independent: contiguous element ranges per GPU, no data-path collective)
    tree, batch = make_batch(args.elems, seed=1000 + rank)
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
    n_active = int(sum(not eng.element_info(int(c), gg_max=2048)["filtered"] for c in ids))
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
    stop.set()
    sampler.join(timeout=2)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    ms_step = float(np.mean(kern_ms))
    cnt = eng.collect() and eng.counters()

    # ---- end to end through the C ABI with host buffers (front-end + H2D + chains + D2H traces + Eval2)
    barrier()
    t0 = time.time()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        eng.load_alignment(Xn, batch.elem_start, batch.elem_len)
        res = eng.run(ids)
    barrier()
    e2e_ms = (time.time() - t0) * 1e3 / e2e_steps
    h2d = int(batch.X.nbytes)
    d2h = int(n_active * 3 * T * (6 * 8 + ((2 * tree.S - 1 + 15) // 16) * 4))
    n_acc = int(np.sum((res.logBF1 > 5) & (res.logBF2 > 5) & (res.filtered == 0)))

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
        hbm_bytes = float(eng_blob_bytes(eng, ids)) + d2h  # alignment streamed once + trace write-back
        # DRAM traffic of the chain kernel per launch: measured once under ncu (newest profiles/r*_v*_traffic.json, bytes per
        # chain-iteration of this workload) and scaled to this launch's chain-iterations; null when the file is absent
        traffic, traffic_file = None, None
        try:
            import glob
            import re

            cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_v*_traffic.json")),
                           key=lambda q: [int(x) for x in re.findall(r"\d+", os.path.basename(q))])
            traffic_file = os.path.basename(cands[-1])
            with open(cands[-1]) as f:
                traffic = float(json.load(f)["dram_bytes_per_chain_iteration"]) * float(cnt["chain_iterations"])
        except Exception:
            pass
        roof = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic, "traffic_source": f"ncu dram__bytes_read+write of one bench-size chain_kernel launch (profiles/{traffic_file})",
                "peak_source": "measured live: register-resident DFMA micro-benchmark (pa_device_fp64_peak)",
                "algorithmic_fma_per_step": fma, "kernel": "chain_kernel",
                "hbm": {"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                        "note": "algorithmic bytes (alignment read once + trace rows written); the kernel's measured DRAM traffic is below"},
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
                "config": {"workload": WORKLOAD, "elements_per_gpu_per_step": args.elems, "active_elements_rank0": n_active,
                           "iterations_per_chain": T, "models": 3, "l2": "flushed between timed steps (256 MiB write)",
                           "parallelism": f"elements sharded over {world} GPU(s), no data-path collective, one all_gather of summaries"},
                "roofline": roof,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "pruned_site_liks": {"value": site_liks, "unit": "full-tree-equivalent site likelihoods/s", "columns_per_pattern": compression},
                "gpu_launches": int(cnt["kernel_launches"]) * args.steps,
                "clocks": _clock_summary(clocks),
                "work": {k: int(v) for k, v in cnt.items()}, "accelerated_calls_rank0": n_acc}
        if not args.no_cpu_baseline:
            base, _ = reference_throughput(tree, os.cpu_count() or 1)
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def eng_blob_bytes(eng, ids):
    # packed alignment bytes read by the chains: 4-bit codes of every pattern for every species
    tot = 0
    for c in ids[: min(len(ids), 256)]:
        info = eng.element_info(int(c), gg_max=2048)
        if not info["filtered"]:
            tot += info["P"] * eng.S // 2
    return tot * (len(ids) / max(1, min(len(ids), 256)))


if __name__ == "__main__":
    main()
