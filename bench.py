#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-site likelihood hot path on B200.

Metric (BASELINE.json): pair-HMM GCUPS and candidate sites/sec.  The JSON line's `value` is the
pair-HMM throughput in GCUPS on the config-3 shape (150 bp reads x 2 haplotype windows of 300 bp,
unbanded, Sum_pairs len_x*len_y / time); the part-2 throughput (AF-grid
likelihood + event integration, sites/s) rides in the same line under "sites_per_sec".

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port on all host cores

A "step" is one pass of the hot path over one synthetic batch (inputs larger than L2, so no flush
is needed between steps).  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  One process per GPU (torchrun); ranks own disjoint shards (region
sharding, no collective on the data path) -> scaling "weak".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_CELL = 14.0  # algorithmic fp64 flops per pair-HMM cell update
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (pairhmm_kernel<5,0,1,0,1,Lin>),
# one launch at the default workload (7168 sites x 140 reads): `ncu --set full` capture summarised in
# profiles/r1_ncu_k1_pairhmm_R5_approx_lut_default_summary.txt (347.3 MB read + 18.6 MB written)
NCU_TRAFFIC_DEFAULT = {"sites": 7168, "reads_per_site": 140, "exact_sum": False, "bytes": 365.9e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", type=int, default=7168, help="candidate sites per step and GPU")
    ap.add_argument("--reads-per-site", type=int, default=140)
    ap.add_argument("--exact-sum", action="store_true",
                    help="time the exact M-state sum instead of bio's ln_sum3_exp_approx")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="0 = auto (about 12 s)")
    ap.add_argument("--cpu-sample-sites", type=int, default=0, help="0 = auto (5-10 s per workload)")
    ap.add_argument("--p2-sites", type=int, default=262144, help="C2 sites per step and GPU (part 2)")
    ap.add_argument("--skip-part2", action="store_true")
    ap.add_argument("--c4-pairs", type=int, default=1184, help="long-read pairs of the C4 subsample (0 = skip)")
    ap.add_argument("--as-reads", type=int, default=262144, help="reads per step for the fused allele-support run")
    ap.add_argument("--part2-only", default="", help="run only this part-2 workload (profiling aid)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def gen_workload(args, rank):
    from libprosic_b200 import synth
    return synth.make_pairhmm_workload(args.sites, args.reads_per_site, seed=synth.SEED_C3 + rank)


def cpu_sample(w, n_pairs, threads, flags):
    """Time the oracle port on the first n_pairs pairs of the workload; returns (GCUPS, seconds)."""
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    from libprosic_b200 import GapParams
    O.lib()
    gap = GapParams().as_ln_array()
    hb, ho, rb, rq, ro = (w["hap_bytes"], w["hap_off"], w["read_bases"], w["read_quals"],
                          w["read_off"])
    ph, pr = w["pair_hap"][:n_pairs], w["pair_read"][:n_pairs]

    def run(k):
        h, r = int(ph[k]), int(pr[k])
        return O.pairhmm_prob_related(hb[ho[h]:ho[h + 1]], rb[ro[r]:ro[r + 1]], rq[ro[r]:ro[r + 1]],
                                      gap, -1, flags)

    t0 = time.perf_counter()
    if threads <= 1:
        for k in range(n_pairs):
            run(k)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL
            list(ex.map(run, range(n_pairs), chunksize=4))
    dt = time.perf_counter() - t0
    cells = sum(int(ho[int(h) + 1] - ho[int(h)]) * int(ro[int(r) + 1] - ro[int(r)])
                for h, r in zip(ph, pr))
    return cells / dt / 1e9, dt, cells


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement (the Rust reference cannot be built in this image) on all
    host cores, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    flags = 0 if args.exact_sum else 1
    small = argparse.Namespace(**vars(args))
    small.sites, small.reads_per_site = 256, 140   # 71 680 pairs of the headline workload's shape
    w = gen_workload(small, 0)
    # calibrate the per-step sample to ~5 s on all cores
    probe = min(len(w["pair_hap"]), cores * 48)
    g, dt, _ = cpu_sample(w, probe, cores, flags)
    per_step = int(max(probe, min(len(w["pair_hap"]), probe * 5.0 / max(dt, 1e-3))))
    for _ in range(max(args.warmup, 0) and 1):
        cpu_sample(w, per_step, cores, flags)
    tot_cells, tot_t = 0, 0.0
    for _ in range(args.steps):
        g, dt, cells = cpu_sample(w, per_step, cores, flags)
        tot_cells += cells
        tot_t += dt
    val = tot_cells / tot_t / 1e9
    line = {"impl": "reference", "metric": "pairhmm_gcups", "value": val, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 pair-HMM: 150bp reads x 2 x 300bp haplotype windows, unbanded",
                       "sample_pairs_per_step": per_step,
                       "m_state_sum": "exact" if args.exact_sum else "ln_sum3_exp_approx"},
            "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores, "kind": "port",
                             "sample": "%d pairs (150x300 cells each) per step on %d threads; oracle "
                                       "restatement, not the Rust binary" % (per_step, cores)},
            "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


FLOPS_PER_EVAL_SINGLE = 34.0  # 10 flops + 1 fp64 log counted as 24
FLOPS_PER_EVAL_CONTAM = 42.0
BYTES_PER_OBS = 74.0          # 9 fp64 columns + 2 bytes of categoricals as laid out in HBM


def part2_workloads(args):
    from libprosic_b200 import scenario
    return {
        "c2_continuous": dict(scn=scenario.single_sample(continuous=True), depths=[30], kind="snv",
                              n_sites=args.p2_sites, af_sets=None, purity=None,
                              desc="C2 scenario B: 1 sample x 30 obs, universe [0,1] on a 31-point "
                                   "Simpson grid, SNV bias configs (1+7)", flops=FLOPS_PER_EVAL_SINGLE),
        "c2_germline": dict(scn=scenario.single_sample(continuous=False), depths=[30], kind="snv",
                            n_sites=args.p2_sites, af_sets=None, purity=None,
                            desc="C2 scenario A: 1 sample x 30 obs, universe {0,.5,1}, SNV bias configs",
                            flops=FLOPS_PER_EVAL_SINGLE),
        "c3_tumor_normal": dict(scn=scenario.tumor_normal(purity=0.75, indel=True), depths=[40, 100],
                                kind="indel", n_sites=max(1024, args.p2_sites // 8),
                                af_sets=[[(0, .6), (.5, .3), (1, .1)], [(0, .4), (.3, .4), (.05, .2)]],
                                purity=[1.0, 0.75],
                                desc="C3 part 2: built-in tumor/normal scenario, 40x normal + 100x tumor, "
                                     "purity .75, tumor grid 101 x normal grid 5, indel bias configs (1+3)",
                                flops=FLOPS_PER_EVAL_CONTAM),
        "c5_trio": dict(scn=scenario.trio(with_biases=True), depths=[30, 30, 30], kind="snv",
                        n_sites=args.p2_sites // 2,
                        af_sets=[[(0, .9), (.5, .07), (1, .03)]] * 3, purity=None,
                        desc="C5 shape: trio (child, father, mother) x 30 obs, genotype universe {0,.5,1} "
                             "per sample, joint events over 27 AF tuples, SNV bias configs, flat prior",
                        flops=FLOPS_PER_EVAL_SINGLE),
    }


def oracle_part2(wl, cols, cat, off, n_sample, threads=1):
    """Oracle on the first n_sample sites with `threads` host threads (POSIX threads inside the C
    oracle, sites dealt round-robin): results, sites/s, seconds."""
    import oracle as O
    scn = wl["scn"]
    spec = O.ModelSpec(scn["n_samples"], scn["resolution"], scn["contaminated"], scn["by"],
                       scn["purity"], scn["nodes"], scn["events"],
                       [O.make_biases(**b) for b in scn["biases"]])
    O.model_compute_batch(spec, cols, cat, off, min(n_sample, 8), 1)  # load / warm
    t0 = time.perf_counter()
    r = O.model_compute_batch(spec, cols, cat, off, n_sample, threads)
    dt = time.perf_counter() - t0
    res = [dict(event_ln=r["event_ln"][s], map_af=r["map_af"][s], n_evals=int(r["n_evals"][s]))
           for s in range(n_sample)]
    return res, n_sample / dt, dt


def executed_utilisation(name):
    """What the kernel EXECUTES next to the survey's flop convention (which prices a software log per
    evaluation the kernel does not take): fp64-pipe and issue-slot utilisation of K3/K4 on this
    workload from the committed ncu summary (profiles/, a separate capture of the same kernel on
    65 536 sites -- never measured under this timed run)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                        "r2_ncu_k3_posterior_%s_summary.txt" % name)
    out = {"executed_fp64_pipe_pct": None, "issue_pct": None, "executed_source": None}
    try:
        for line in open(path):
            f = line.split()
            if len(f) >= 2 and f[0] == "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active":
                out["executed_fp64_pipe_pct"] = float(f[1])
            if len(f) >= 2 and f[0] == "sm__inst_issued.avg.pct_of_peak_sustained_active":
                out["issue_pct"] = float(f[1])
        out["executed_source"] = "profiles/" + os.path.basename(path)
    except OSError:
        pass
    return out


def bench_part2(args, ctx, torch, dist, dev, stream, rank, world, name, wl, timed):
    """sites/s of AF-grid likelihood + event integration (K3/K4) on one synthetic shard per rank."""
    from libprosic_b200 import synth
    scn = wl["scn"]
    ns, ne = scn["n_samples"], len(scn["events"])
    cols, cat, off, _ = synth.make_pileups(wl["n_sites"], wl["depths"], seed=synth.SEED_C2 + 17 * rank,
                                           kind=wl["kind"], af_sets=wl["af_sets"], purity=wl["purity"],
                                           artifact_frac=0.05)
    n_sites = wl["n_sites"]
    names = synth.OBS_COLUMNS
    h_cols = {n: torch.from_numpy(np.ascontiguousarray(cols[n])).pin_memory() for n in names}
    h_cat = torch.from_numpy(cat.view(np.int16)).pin_memory()
    h_off = torch.from_numpy(off).pin_memory()
    d_cols = {n: h_cols[n].to(dev) for n in names}
    d_cat, d_off = h_cat.to(dev), h_off.to(dev)
    d_ev = torch.empty(n_sites * ne, dtype=torch.float64, device=dev)
    d_marg = torch.empty(n_sites, dtype=torch.float64, device=dev)
    d_af = torch.empty(n_sites * ns, dtype=torch.float64, device=dev)
    d_mb = torch.empty(n_sites * ns, dtype=torch.int32, device=dev)
    ptrs = {n: d_cols[n].data_ptr() for n in names}

    def step_dev():
        ctx.posterior_batch_dev(scn, n_sites, ptrs, d_cat.data_ptr(), d_off.data_ptr(), max(wl["depths"]),
                                d_ev.data_ptr(), d_marg.data_ptr(), d_af.data_ptr(), d_mb.data_ptr())

    np_cols = {n: h_cols[n].numpy() for n in names}
    np_cat = h_cat.numpy().view(np.uint16)

    def pinned_out():
        return dict(event_ln=torch.empty((n_sites, ne), dtype=torch.float64).pin_memory().numpy(),
                    marginal=torch.empty(n_sites, dtype=torch.float64).pin_memory().numpy(),
                    map_af=torch.empty((n_sites, ns), dtype=torch.float64).pin_memory().numpy(),
                    map_bias=torch.empty((n_sites, ns), dtype=torch.int32).pin_memory().numpy())
    e2e_out, f32_out = pinned_out(), pinned_out()

    def step_e2e():
        return ctx.posterior_batch(scn, np_cols, np_cat, h_off.numpy(), out=e2e_out)

    # compact entry: the seven stored fields as f32 (lossless for MiniLogProb-quantised values)
    stored = ("prob_mapping", "prob_alt", "prob_ref", "prob_missed_allele", "prob_sample_alt",
              "prob_double_overlap", "prob_hit_base")
    h_cols32 = {n: torch.from_numpy(np.ascontiguousarray(cols[n], dtype=np.float32)).pin_memory() for n in stored}
    np_cols32 = {n: h_cols32[n].numpy() for n in stored}

    def step_e2e_f32():
        return ctx.posterior_batch_f32(scn, np_cols32, np_cat, h_off.numpy(), out=f32_out)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    l0 = ctx.launch_count()
    ms = timed(step_dev, args.steps)
    launches = ctx.launch_count() - l0
    out_e2e = step_e2e()
    t0 = time.perf_counter()
    timed(step_e2e, args.steps)
    ms_e2e = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    out32 = step_e2e_f32()
    t0 = time.perf_counter()
    timed(step_e2e_f32, args.steps)
    ms_e2e32 = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms_e2e32], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e32 = float(t.item())
    # the records as `preprocess` stores them (bincode words per field), decoded on the device
    from libprosic_b200 import obs_codec as lc
    wire = lc.encode_batch(cols, cat, off, pin=True)
    wire_out = pinned_out()

    def step_e2e_wire():
        return ctx.posterior_batch_wire(scn, wire, wire_out)

    out_wire = {k: v.copy() for k, v in step_e2e_wire().items()}
    t0 = time.perf_counter()
    timed(step_e2e_wire, args.steps)
    ms_e2ew = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms_e2ew], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2ew = float(t.item())
    in_bytes32 = sum(int(v.numel() * 4) for v in h_cols32.values()) + int(h_cat.numel() * 2 + h_off.numel() * 8)
    in_bytes = sum(int(v.numel() * v.element_size()) for v in h_cols.values()) + \
        int(h_cat.numel() * 2 + h_off.numel() * 8)
    out_bytes = n_sites * (ne * 8 + 8 + ns * 12)
    res = {"workload": wl["desc"], "sites_per_gpu": n_sites, "value": n_sites * world * args.steps / (ms * 1e-3),
           "unit": "sites/s", "ms_per_step": ms / args.steps, "gpu_launches": int(launches),
           "e2e": {"value": n_sites * world * args.steps / (ms_e2e * 1e-3), "unit": "sites/s",
                   "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                   "ms_per_step": ms_e2e / args.steps,
                   "host_gb_per_s": float(in_bytes + out_bytes) / (ms_e2e / args.steps * 1e-3) / 1e9},
           "e2e_f32": {"value": n_sites * world * args.steps / (ms_e2e32 * 1e-3), "unit": "sites/s",
                       "h2d_bytes_per_step": in_bytes32, "d2h_bytes_per_step": out_bytes,
                       "ms_per_step": ms_e2e32 / args.steps,
                       "host_gb_per_s": float(in_bytes32 + out_bytes) / (ms_e2e32 / args.steps * 1e-3) / 1e9,
                       "note": "vlr_posterior_batch_f32: stored fields as f32, derived columns on the device"},
           "e2e_wire": {"value": n_sites * world * args.steps / (ms_e2ew * 1e-3), "unit": "sites/s",
                        "h2d_bytes_per_step": int(wire.nbytes), "d2h_bytes_per_step": out_bytes,
                        "ms_per_step": ms_e2ew / args.steps,
                        "wire_bytes_per_observation": float(wire.nbytes) / max(1, len(cat)),
                        "host_gb_per_s": float(wire.nbytes + out_bytes) / (ms_e2ew / args.steps * 1e-3) / 1e9,
                        "note": "vlr_posterior_batch_wire: the records as `preprocess` stores them (bincode "
                                "words per field, format 8) uploaded as they are and decoded by a kernel"}}
    if rank == 0:
        # the derived columns (prob_mismapping, prob_single_overlap) come from the device's libm here and
        # from numpy in the f64 entry's inputs: last-bit differences in ln space, identical MAP calls
        res["e2e_wire"]["max_abs_ln_vs_f64_entry"] = float(np.nanmax(np.abs(np.where(
            np.isfinite(out_wire["event_ln"]) & np.isfinite(out_e2e["event_ln"]),
            out_wire["event_ln"] - out_e2e["event_ln"], 0.0))))
        res["e2e_wire"]["map_af_identical_to_f64_entry"] = bool(
            np.array_equal(out_wire["map_af"], out_e2e["map_af"], equal_nan=True))
        res["e2e_f32"]["max_abs_ln_vs_f64_entry"] = float(np.nanmax(np.abs(np.where(
            np.isfinite(out32["event_ln"]) & np.isfinite(out_e2e["event_ln"]),
            out32["event_ln"] - out_e2e["event_ln"], 0.0))))
        n_cpu = min(n_sites, args.cpu_sample_sites or (16384 if ns == 1 else (4096 if ns == 3 else 256)))
        ores, cpu_sps, cpu_dt = oracle_part2(wl, cols, cat, off, n_cpu)
        ev = d_ev.cpu().numpy().reshape(n_sites, ne)[:n_cpu]
        want = np.array([r["event_ln"] for r in ores])
        fin = np.isfinite(ev) & np.isfinite(want)
        res["parity_max_abs_ln"] = float(np.max(np.abs(ev[fin] - want[fin]))) if fin.any() else 0.0
        res["parity_map_af_identical"] = bool(np.array_equal(
            d_af.cpu().numpy().reshape(n_sites, ns)[:n_cpu], np.array([r["map_af"] for r in ores]),
            equal_nan=True))
        res["e2e"]["identical_to_device_path"] = bool(np.array_equal(
            out_e2e["event_ln"], d_ev.cpu().numpy().reshape(n_sites, ne), equal_nan=True))
        n_obs_site = float(np.sum(wl["depths"]))
        # distinct likelihood evaluations per site (cache-miss semantics of likelihood.rs:128,221):
        # oracle count of distinct keys x observations of the sample they run over
        keys = float(np.mean([r["n_evals"] for r in ores]))
        evals = keys * (wl["depths"][-1] if ns == 2 else wl["depths"][0])
        sps_gpu = n_sites * args.steps / (ms * 1e-3)
        res["evals_per_site"] = evals
        res["roofline"] = {"bound": "fp64", "achieved": sps_gpu * evals * wl["flops"] / 1e12,
                           "unit": "TFLOP/s", "flops_per_eval": wl["flops"],
                           "note": "survey convention (a software fp64 log = 24 flops per evaluation); the "
                                   "kernel takes one log per key instead of one per observation",
                           "hbm": {"achieved": sps_gpu * n_obs_site * BYTES_PER_OBS / 1e9, "unit": "GB/s"}}
        res["roofline"].update(executed_utilisation(name))
        res["cpu_baseline"] = {"value": cpu_sps, "unit": "sites/s", "cores": 1, "kind": "port",
                               "sample": "%d sites of this workload, 1 thread, %.1f s" % (n_cpu, cpu_dt)}
        cores = os.cpu_count() or 1
        n_all = min(n_sites, n_cpu * min(cores, 8))
        _, all_sps, all_dt = oracle_part2(wl, cols, cat, off, n_all, cores)
        res["cpu_baseline_all_cores"] = {"value": all_sps, "unit": "sites/s", "cores": cores, "kind": "port",
                                         "sample": "%d sites of this workload, %d threads, %.1f s" % (n_all, cores, all_dt)}
    return res


def bench_allele_support(args, ctx, w, flags, rank, world, torch, dist, dev):
    """Part 1 as the pipeline runs it (Realigner::allele_support per read): Myers best hits,
    minimal-distance selection, certainty short-circuit, banded pair-HMM on the shrunk windows,
    normalisation -- one fused device pipeline behind vlr_allele_support_batch (host buffers,
    end to end)."""
    import libprosic_b200 as P
    n_reads = min(len(w["read_off"]) - 1, args.as_reads)
    site_of = w["site_of"][:n_reads]
    region_read = np.arange(n_reads, dtype=np.int32)
    region_haps = (2 * site_of[:, None] + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    region_hap_off = np.arange(0, 2 * n_reads + 1, 2, dtype=np.int32)
    region_n_ref = np.ones(n_reads, dtype=np.int32)
    gap = P.GapParams()

    # inputs and outputs in pinned host memory, as a host integration would hold its batch buffers
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    p_in = {k: pin(v) for k, v in dict(rr=region_read, rho=region_hap_off, rh=region_haps, rnr=region_n_ref,
                                      hb=w["hap_bytes"], ho=w["hap_off"], rb=w["read_bases"][:int(w["read_off"][n_reads])],
                                      rq=w["read_quals"][:int(w["read_off"][n_reads])], ro=w["read_off"][:n_reads + 1]).items()}
    p_out = dict(prob_ref=pin(np.empty(n_reads)), prob_alt=pin(np.empty(n_reads)), raw_ref=pin(np.empty(n_reads)),
                 raw_alt=pin(np.empty(n_reads)), ref_dist=pin(np.empty(n_reads, np.int32)),
                 alt_dist=pin(np.empty(n_reads, np.int32)), n_indel_ops=pin(np.empty(n_reads, np.int32)),
                 indel_ops=pin(np.zeros((n_reads, 32), np.uint8)))

    def run():
        return ctx.allele_support_batch(p_in["rr"], p_in["rho"], p_in["rh"], p_in["rnr"], p_in["hb"], p_in["ho"],
                                        p_in["rb"], p_in["rq"], p_in["ro"], gap, flags=flags, out=p_out)
    out = run()
    run()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = run()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    cells = 2.0 * n_reads * 150 * 300
    res = {"workload": "C3 reads: Realigner::allele_support per read (ref + alt haplotype, band dist+2)",
           "reads_per_gpu": int(n_reads), "value": n_reads * world * args.steps / dt, "unit": "reads/s",
           "effective_gcups_full_matrix": cells * world * args.steps / dt / 1e9,
           "ms_per_step": 1e3 * dt / args.steps, "timing": "host wall clock around the synchronous C-ABI call",
           "frac_pairs_certainty_shortcut": float(np.mean(np.concatenate([out["ref_dist"], out["alt_dist"]]) == 0))}
    # K batches in flight: K host threads, each with its own context (stream, scratch) and 1/K of the
    # reads per step -- the copies and host checks of one batch run under the kernels of the others
    # ("one host thread + 2-3 streams + pinned double buffers per GPU")
    K = 4
    ctxs = [P.Context(dev.index) for _ in range(K)]
    bnd = [n_reads * i // K for i in range(K + 1)]
    jobs = []
    for i in range(K):
        r0, r1 = bnd[i], bnd[i + 1]
        b0, b1 = int(w["read_off"][r0]), int(w["read_off"][r1])
        so = site_of[r0:r1]
        s0, s1 = int(so[0]), int(so[-1]) + 1
        h0, h1 = int(w["hap_off"][2 * s0]), int(w["hap_off"][2 * s1])
        jobs.append((dict(rr=pin(np.arange(r1 - r0, dtype=np.int32)), rho=pin(np.arange(0, 2 * (r1 - r0) + 1, 2, dtype=np.int32)),
                          rh=pin((2 * (so - s0)[:, None] + np.arange(2)[None, :]).reshape(-1).astype(np.int32)),
                          rnr=pin(np.ones(r1 - r0, dtype=np.int32)), hb=pin(w["hap_bytes"][h0:h1]),
                          ho=pin(w["hap_off"][2 * s0:2 * s1 + 1] - h0), rb=pin(w["read_bases"][b0:b1]),
                          rq=pin(w["read_quals"][b0:b1]), ro=pin(w["read_off"][r0:r1 + 1] - b0)),
                     {k: v[r0:r1] for k, v in p_out.items()}))
    out_single = {k: v.copy() for k, v in out.items()}

    def work(i, steps):
        ji, jo = jobs[i]
        for _ in range(steps):
            ctxs[i].allele_support_batch(ji["rr"], ji["rho"], ji["rh"], ji["rnr"], ji["hb"], ji["ho"], ji["rb"],
                                         ji["rq"], ji["ro"], gap, flags=flags, out=jo)

    def run_k(steps):
        ts = [threading.Thread(target=work, args=(i, steps)) for i in range(K)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    run_k(2)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run_k(args.steps)
    dtk = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dtk], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dtk = float(t.item())
    res["batches_in_flight_4"] = {
        "value": n_reads * world * args.steps / dtk, "unit": "reads/s", "ms_per_step": 1e3 * dtk / args.steps,
        "note": "4 host threads x 4 contexts, a quarter of the reads each per step",
        "identical_to_single_call": bool(all(np.array_equal(out_single[k], p_out[k], equal_nan=True) for k in p_out))}
    del ctxs
    if rank == 0:
        import oracle as O
        n_cpu = min(n_reads, 4096)
        t0 = time.perf_counter()
        worst = 0.0
        for r in range(n_cpu):
            ids = region_haps[2 * r:2 * r + 2]
            hp = [w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]] for h in ids]
            want = O.allele_support_region([hp[0]], [hp[1]], w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
                                           w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]],
                                           gap.as_ln_array(), flags=flags)
            worst = max(worst, abs(want["prob_ref"] - out_single["prob_ref"][r]),
                        abs(want["prob_alt"] - out_single["prob_alt"][r]))
        cdt = time.perf_counter() - t0
        res["parity_max_abs_ln"] = worst
        res["cpu_baseline"] = {"value": n_cpu / cdt, "unit": "reads/s", "cores": 1, "kind": "port",
                               "sample": "%d reads, 1 thread, %.1f s (oracle: DP-matrix edit distance + log-space banded HMM)" % (n_cpu, cdt)}
        # all host cores (ctypes releases the GIL around the C oracle; ~0.5 ms of C per call)
        from concurrent.futures import ThreadPoolExecutor
        cores = os.cpu_count() or 1
        n_all = min(n_reads, n_cpu * min(cores, 8))
        gl = gap.as_ln_array()

        def one(r):
            ids = region_haps[2 * r:2 * r + 2]
            hp = [w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]] for h in ids]
            O.allele_support_region([hp[0]], [hp[1]], w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
                                    w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gl, flags=flags)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(one, range(n_all), chunksize=16))
        adt = time.perf_counter() - t0
        res["cpu_baseline_all_cores"] = {"value": n_all / adt, "unit": "reads/s", "cores": cores, "kind": "port",
                                         "sample": "%d reads, %d threads, %.1f s" % (n_all, cores, adt)}
    return res


def bench_c1(ctx):
    """Config C1 (the reference's bundled tumor/normal testcases, tests/golden/tn): preprocess + call
    end to end with both compute stages on the GPU; counts the testcases whose `expected:` block
    (the reference's own integration-test assertions) holds."""
    import glob
    from libprosic_b200 import fixtures as F
    paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "tn", "*.json.gz")))
    be = F.GpuBackend(ctx)
    t0 = time.perf_counter()
    held, failed, n_obs = 0, [], 0
    not_restated = []
    for p in paths:
        try:
            r = F.run_testcase(p, be)
        except NotImplementedError:   # host glue the fixture runner does not restate
            not_restated.append(os.path.basename(p).split(".")[0])
            continue
        n_obs += r["n_obs"]["normal"] + r["n_obs"]["tumor"]
        if all(v for _, v in F.check_expectations(r)):
            held += 1
        else:
            failed.append(os.path.basename(p).split(".")[0])
    # Generic-mode testcases driven by their own scenario.yaml through the grammar front end
    spaths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scn", "*.json.gz")) +
                    glob.glob(os.path.join(ROOT, "tests", "golden", "scn_prior", "*.json.gz")))
    s_held, s_failed, s_asserted = 0, [], 0
    for p in spaths:
        r = F.run_scenario_testcase(p, be)
        n_obs += sum(r["n_obs"].values())
        v = F.check_expectations(r)
        s_asserted += bool(v)
        if all(x for _, x in v):
            s_held += bool(v)
        else:
            s_failed.append(os.path.basename(p).split(".")[0])
    return {"workload": "C1: %d tumor/normal indel testcases and %d Generic (scenario.yaml) testcases of the "
                        "reference (real reads), preprocess + call through the restated host glue with K2/K1 and "
                        "K3/K4 on the GPU" % (len(paths), len(spaths)),
            "testcases": len(paths) - len(not_restated), "expectations_hold": held, "not_holding": failed,
            "not_restated": not_restated,
            "scenario_testcases": len(spaths), "scenario_with_expectations": s_asserted,
            "scenario_expectations_hold": s_held, "scenario_not_holding": s_failed,
            "observations": n_obs, "seconds": time.perf_counter() - t0,
            "note": "test21 is the documented exception (DESIGN.md section 5); all 32 tumor/normal testcases: "
                    "tests/golden/tn_results.txt (31 hold); all runnable Generic testcases: "
                    "tests/golden/scenario_results.txt (22 of 25 hold = all that upstream asserts)"}


def bench_c4(args, ctx, torch, dist, dev, rank, world, timed):
    """Config C4 (long reads 15 kb x 20 kb haplotype windows), labelled SUBSAMPLE of the 1 M pairs:
    strip-wise K1 with per-strip renormalisation; gap parameters raised as in SURVEY 8(d)."""
    import libprosic_b200 as P
    from libprosic_b200 import synth
    n = args.c4_pairs
    w = synth.make_long_read_pairs(n, seed=synth.SEED_C4 + rank)
    # one small pair appended for the oracle spot check (a 15 kb x 20 kb pair takes ~25 s on the CPU)
    small = synth.make_long_read_pairs(1, seed=7, read_len=1500, hap_len=2200)
    hap = np.concatenate([w["hap_bytes"], small["hap_bytes"]])
    hoff = np.concatenate([w["hap_off"], w["hap_off"][-1:] + small["hap_off"][1:]])
    rb = np.concatenate([w["read_bases"], small["read_bases"]])
    rq = np.concatenate([w["read_quals"], small["read_quals"]])
    roff = np.concatenate([w["read_off"], w["read_off"][-1:] + small["read_off"][1:]])
    gap = P.GapParams(0.02, 0.02, 0.3, 0.3)
    npairs = n + 1
    d_hap = torch.zeros(len(hap) + 64, dtype=torch.uint8, device=dev)
    d_hap[:len(hap)] = torch.from_numpy(hap).to(dev)
    d_hoff, d_roff = torch.from_numpy(hoff).to(dev), torch.from_numpy(roff).to(dev)
    d_rb, d_rq = torch.from_numpy(rb).to(dev), torch.from_numpy(rq).to(dev)
    ws_bytes = ctx.pairhmm_workspace_bytes(npairs, 20000, 15000)
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(npairs, dtype=torch.float64, device=dev)

    def step():
        ctx.pairhmm_batch_dev(npairs, d_hap.data_ptr(), d_hoff.data_ptr(), npairs, d_rb.data_ptr(),
                              d_rq.data_ptr(), d_roff.data_ptr(), npairs, 0, 0, 0, gap, P.HMM_DEFAULT,
                              d_ws.data_ptr(), ws_bytes, d_out.data_ptr())
    step()
    ms = timed(step, 1)
    cells = w["cells"] + small["cells"]
    res = {"workload": "C4 long reads: 15 kb reads x 20 kb haplotype windows, gap open .02 / extend .3, "
                       "SUBSAMPLE of %d of the 1 M pairs per GPU" % n,
           "pairs_per_gpu": n, "value": cells * world / (ms * 1e-3) / 1e9, "unit": "GCUPS",
           "ms_per_step": ms, "m_state_sum": "ln_sum3_exp_approx (reference)"}
    if rank == 0:
        import oracle as O
        got = float(d_out[n].item())
        want = O.pairhmm_prob_related(small["hap_bytes"], small["read_bases"], small["read_quals"],
                                      gap.as_ln_array(), -1, P.HMM_DEFAULT)
        res["parity_abs_ln_on_1500x2200_pair"] = abs(got - want)
        out = d_out[:n].cpu().numpy()
        res["ln_prob_range"] = [float(out.min()), float(out.max())]
        res["finite"] = bool(np.isfinite(out).all())
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import libprosic_b200 as P
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_binding = None
    if world > 1:
        # one process per GPU: run on (and first-touch the pinned buffers on) the GPU's own NUMA node.
        # The CPU legs only run at N=1, where the process keeps all host cores.
        from libprosic_b200.sharding import bind_host_to_device
        pr = torch.cuda.get_device_properties(local_rank)
        bus = ("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
               if hasattr(pr, "pci_bus_id") else None)
        host_binding = bind_host_to_device(local_rank, bus)
        dist.init_process_group("nccl", device_id=dev)
    ctx = P.Context(local_rank)
    # a real (non-default) stream: the context enqueues on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    gap = P.GapParams()
    flags = 0 if args.exact_sum else P.HMM_DEFAULT

    # ---- host <-> device copy rate of this box (pinned memory, one 256 MB copy, best of 5): the roof of
    # every host-buffer ("e2e") leg whose kernels are faster than its copies
    def copy_rate(h2d):
        nbytes = 256 << 20
        hbuf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if h2d:
                dbuf.copy_(hbuf, non_blocking=True)
            else:
                hbuf.copy_(dbuf, non_blocking=True)
            e1.record(stream)
            e1.synchronize()
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        return best
    pcie = {"h2d_gb_per_s": copy_rate(True), "d2h_gb_per_s": copy_rate(False),
            "how": "one 256 MB pinned copy on the bench stream, best of 5, CUDA events"}

    # ---- synthetic shard of this rank (config C3 shape), pinned on the host for the e2e path
    w = gen_workload(args, rank)
    n_pairs = len(w["pair_hap"])
    n_haps, n_reads = len(w["hap_off"]) - 1, len(w["read_off"]) - 1
    cells = w["cells"]
    host = {k: torch.from_numpy(w[k]).pin_memory() for k in
            ("hap_bytes", "hap_off", "read_bases", "read_quals", "read_off", "pair_hap", "pair_read")}
    h_out = torch.empty(n_pairs, dtype=torch.float64).pin_memory()
    # device-resident copies for the kernel-only number
    d = {k: v.to(dev) for k, v in host.items()}
    d_hap = torch.zeros(d["hap_bytes"].numel() + 64, dtype=torch.uint8, device=dev)
    d_hap[:d["hap_bytes"].numel()] = d["hap_bytes"]
    ws_bytes = ctx.pairhmm_workspace_bytes(n_pairs)
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_pairs, dtype=torch.float64, device=dev)

    def step_dev(fl=None):
        ctx.pairhmm_batch_dev(n_pairs, d_hap.data_ptr(), d["hap_off"].data_ptr(), n_haps,
                              d["read_bases"].data_ptr(), d["read_quals"].data_ptr(),
                              d["read_off"].data_ptr(), n_reads, d["pair_hap"].data_ptr(),
                              d["pair_read"].data_ptr(), 0, gap, flags if fl is None else fl,
                              d_ws.data_ptr(), ws_bytes, d_out.data_ptr())

    def step_e2e():
        ctx.pairhmm_batch(host["hap_bytes"].numpy(), host["hap_off"].numpy(),
                          host["read_bases"].numpy(), host["read_quals"].numpy(),
                          host["read_off"].numpy(), gap, host["pair_hap"].numpy(),
                          host["pair_read"].numpy(), flags=flags, out=h_out.numpy())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms = timed(step_dev, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # the other M-state-sum mode, for the record (same inputs, same timing discipline)
    other = P.HMM_DEFAULT if args.exact_sum else 0
    for _ in range(2):
        step_dev(other)
    ms_other = timed(lambda: step_dev(other), args.steps)
    step_dev()  # leave the headline mode's results in d_out for the parity spot check

    # ---- end-to-end through the host-buffer C-ABI call (H2D + kernels + D2H inside the timed region)
    step_e2e()
    t0 = time.perf_counter()
    ms_e2e_ev = timed(step_e2e, args.steps)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e_ev, 0.0)
    if world > 1:
        t = torch.tensor([wall_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_e2e = float(t.item())
    # the host call is synchronous: its cost is wall time, CUDA events on the stream agree to the ms
    ms_e2e = max(ms_e2e, wall_e2e) if ms_e2e < 0.5 * wall_e2e else ms_e2e

    # ---- parity spot check against the oracle on a few pairs (outside the timed regions)
    parity = None
    if rank == 0:
        import oracle as O
        got = d_out[:64].cpu().numpy()
        want = np.array([O.pairhmm_prob_related(
            w["hap_bytes"][w["hap_off"][h]:w["hap_off"][h + 1]],
            w["read_bases"][w["read_off"][r]:w["read_off"][r + 1]],
            w["read_quals"][w["read_off"][r]:w["read_off"][r + 1]], gap.as_ln_array(), -1, flags)
            for h, r in zip(w["pair_hap"][:64], w["pair_read"][:64])])
        parity = float(np.max(np.abs(got - want)))
        same_e2e = bool(np.array_equal(h_out.numpy(), d_out.cpu().numpy()))

    allele_support = None
    if not args.skip_part2:
        allele_support = bench_allele_support(args, ctx, w, flags, rank, world, torch, dist, dev)
    c1 = None
    if not args.skip_part2 and rank == 0 and not args.part2_only:
        c1 = bench_c1(ctx)
    c4 = None
    if not args.skip_part2 and args.c4_pairs > 0:
        del d, d_hap, d_ws, d_out
        torch.cuda.empty_cache()
        c4 = bench_c4(args, ctx, torch, dist, dev, rank, world, timed)
    part2 = {}
    if not args.skip_part2:
        torch.cuda.empty_cache()
        for name, wl in part2_workloads(args).items():
            if args.part2_only and name != args.part2_only:
                continue
            part2[name] = bench_part2(args, ctx, torch, dist, dev, stream, rank, world, name, wl, timed)

    total_cells = cells * world
    gcups = total_cells * args.steps / (ms * 1e-3) / 1e9
    gcups_e2e = total_cells * args.steps / (ms_e2e * 1e-3) / 1e9
    if rank == 0:
        peak_tf = ctx.measure_fp64_peak()
        ach_tf = cells * args.steps * FLOPS_PER_CELL / (ms * 1e-3) / 1e12  # this rank's GPU
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        in_bytes = sum(int(host[k].numel() * host[k].element_size()) for k in host)
        cores = os.cpu_count() or 1
        # bounded CPU sample of the same workload, one thread (the reference is single-threaded)
        n_cpu = min(n_pairs, args.cpu_sample_pairs or 8192)
        g1, dt1, _ = cpu_sample(w, n_cpu, 1, flags)
        line = {
            "metric": "pairhmm_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C3 pair-HMM: 150bp reads x 2 x 300bp haplotype windows, unbanded",
                       "sites_per_gpu": args.sites, "reads_per_site": args.reads_per_site,
                       "pairs_per_gpu": n_pairs, "cells_per_step_per_gpu": cells,
                       "m_state_sum": "exact" if args.exact_sum else "ln_sum3_exp_approx (reference)",
                       "gap_params": "cli defaults 2.8e-6/5.1e-6/0/0", "parallelism": "region-sharded x%d" % world,
                       "host_binding": host_binding,
                       "l2": "inputs %.0f MB per step > 126 MB L2, no flush" % (in_bytes / 1e6)},
            "e2e": {"value": gcups_e2e, "unit": "GCUPS", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": n_pairs * 8, "ms_per_step": ms_e2e / args.steps,
                    "identical_to_device_path": same_e2e},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": ach_tf / peak_tf if peak_tf else None,
                         "traffic": (NCU_TRAFFIC_DEFAULT["bytes"]
                                     if (args.sites, args.reads_per_site, args.exact_sum) ==
                                     (NCU_TRAFFIC_DEFAULT["sites"], NCU_TRAFFIC_DEFAULT["reads_per_site"],
                                      NCU_TRAFFIC_DEFAULT["exact_sum"]) else None),
                         "traffic_unit": "bytes per launch of the dominant kernel (ncu, profiles/)",
                         "algorithmic_bytes_per_launch": in_bytes + n_pairs * 8,
                         "bound_note": "the path is bound by the fp64 pipe, not by HBM or tensor cores "
                                       "(SURVEY.md 8d): 0.004 B/cell; the HBM figure is under 'hbm'",
                         "peak_source": "vlr_measure_fp64_peak (DFMA chains, measured live); "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "flops_per_cell": FLOPS_PER_CELL,
                         "hbm": {"achieved": in_bytes * args.steps / (ms * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s"}},
            "cpu_baseline": {"value": g1, "unit": "GCUPS", "cores": 1, "kind": "port",
                             "sample": "%d pairs of this workload, 1 thread of %d host cores, %.1f s; "
                                       "oracle restatement (C), not the Rust binary" % (n_cpu, cores, dt1)},
            "other_mode": {"m_state_sum": "ln_sum3_exp_approx" if args.exact_sum else "exact",
                           "value": total_cells * args.steps / (ms_other * 1e-3) / 1e9, "unit": "GCUPS"},
            "parity_max_abs_ln": parity,
            "allele_support": allele_support,
            "c1_testcases": c1,
            "c4_long_reads": c4,
            "sites_per_sec": part2,
            "host_copy_peak": pcie,
            "clocks": clocks,
        }
        for v in part2.values():
            if "roofline" in v:
                v["roofline"]["peak"] = peak_tf
                v["roofline"]["frac"] = v["roofline"]["achieved"] / peak_tf if peak_tf else None
                v["roofline"]["hbm"]["peak"] = hbm_peak
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
