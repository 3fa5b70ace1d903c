#!/usr/bin/env python
"""Host-buffer (e2e) timing of the headline pair-HMM batch of bench.py, for tuning the copy pipeline
(VLR_K1_CHUNK_MB = size of the first chunk).

    python tools/e2e_probe.py [--steps K]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import bench
    import libprosic_b200 as P
    args = argparse.Namespace(sites=7168, reads_per_site=140)
    w = bench.gen_workload(args, 0)
    ctx = P.Context(0)
    host = {k: torch.from_numpy(w[k]).pin_memory().numpy() for k in
            ("hap_bytes", "hap_off", "read_bases", "read_quals", "read_off", "pair_hap", "pair_read")}
    out = torch.empty(len(w["pair_hap"]), dtype=torch.float64).pin_memory().numpy()
    gap = P.GapParams()

    def step():
        ctx.pairhmm_batch(host["hap_bytes"], host["hap_off"], host["read_bases"], host["read_quals"], host["read_off"],
                          gap, host["pair_hap"], host["pair_read"], flags=P.HMM_DEFAULT, out=out)
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = (time.perf_counter() - t0) / a.steps
    print("VLR_K1_CHUNK_MB=%s: %.2f ms per step, %.1f GCUPS end to end" % (os.environ.get("VLR_K1_CHUNK_MB", "default"), dt * 1e3, w["cells"] / dt / 1e9))


if __name__ == "__main__":
    main()
