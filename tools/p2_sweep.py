#!/usr/bin/env python
"""Device-resident timing of the part-2 workloads of bench.py (K3/K4), for tuning sweeps.

    python tools/p2_sweep.py [workload ...] [--sites N] [--steps K] [--check N]

Prints one JSON line per workload: sites/s (CUDA events on the context's stream, after warm-up)
and the maximum |delta ln| against the oracle on the first `--check` sites.  Tuning knobs are read
from the environment by the library (e.g. VLR_POST_MINB).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*")
    ap.add_argument("--sites", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=512)
    ap.add_argument("--env", default="", help="NAME=v1,v2,...: repeat the timing with each value of the knob")
    a = ap.parse_args()
    import torch
    import bench
    import libprosic_b200 as P
    from libprosic_b200 import synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = P.Context(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    args = argparse.Namespace(p2_sites=a.sites)
    wls = bench.part2_workloads(args)
    for name in (a.workloads or list(wls)):
        wl = wls[name]
        scn = wl["scn"]
        ns, ne = scn["n_samples"], len(scn["events"])
        n_sites = wl["n_sites"]
        cols, cat, off, _ = synth.make_pileups(n_sites, wl["depths"], seed=synth.SEED_C2, kind=wl["kind"],
                                               af_sets=wl["af_sets"], purity=wl["purity"], artifact_frac=0.05)
        d_cols = {n: torch.from_numpy(np.ascontiguousarray(cols[n])).to(dev) for n in synth.OBS_COLUMNS}
        d_cat = torch.from_numpy(cat.view(np.int16)).to(dev)
        d_off = torch.from_numpy(off).to(dev)
        d_ev = torch.empty(n_sites * ne, dtype=torch.float64, device=dev)
        d_marg = torch.empty(n_sites, dtype=torch.float64, device=dev)
        d_af = torch.empty(n_sites * ns, dtype=torch.float64, device=dev)
        d_mb = torch.empty(n_sites * ns, dtype=torch.int32, device=dev)
        ptrs = {n: d_cols[n].data_ptr() for n in synth.OBS_COLUMNS}

        def step():
            ctx.posterior_batch_dev(scn, n_sites, ptrs, d_cat.data_ptr(), d_off.data_ptr(), max(wl["depths"]),
                                    d_ev.data_ptr(), d_marg.data_ptr(), d_af.data_ptr(), d_mb.data_ptr())
        # --env "A=1,2;B=0,1" sweeps the cartesian product of the knobs
        import itertools
        knobs = [(kv.split("=")[0], kv.split("=")[1].split(",")) for kv in a.env.split(";")] if a.env else []
        vals = list(itertools.product(*[v for _, v in knobs])) or [None]
        for val in vals:
          for (k, _), v in zip(knobs, val or ()):
              os.environ[k] = v
          for _ in range(3):
              step()
          torch.cuda.synchronize(dev)
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record(stream)
          for _ in range(a.steps):
              step()
          e1.record(stream)
          e1.synchronize()
          ms = e0.elapsed_time(e1) / a.steps
          res = {"workload": name, "sites": n_sites, "ms_per_step": ms, "sites_per_s": n_sites / (ms * 1e-3),
                 "env": {k: v for k, v in os.environ.items() if k.startswith("VLR_")}}
          if val is not vals[-1]:
              print(json.dumps(res), flush=True)
        if a.check:
            n = min(a.check, n_sites)
            ores, _, _ = bench.oracle_part2(wl, cols, cat, off, n)
            ev = d_ev.cpu().numpy().reshape(n_sites, ne)[:n]
            want = np.array([r["event_ln"] for r in ores])
            fin = np.isfinite(ev) & np.isfinite(want)
            res["max_abs_ln"] = float(np.max(np.abs(ev[fin] - want[fin]))) if fin.any() else 0.0
            res["nonfinite_same"] = bool(np.array_equal(np.isfinite(ev), np.isfinite(want)))
            res["map_af_identical"] = bool(np.array_equal(d_af.cpu().numpy().reshape(n_sites, ns)[:n],
                                                          np.array([r["map_af"] for r in ores]), equal_nan=True))
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
