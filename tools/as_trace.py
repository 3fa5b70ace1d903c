#!/usr/bin/env python
"""Lap timers of vlr_allele_support_batch (VLR_TRACE) on the C3 read set of bench.py.

    python tools/as_trace.py [--reads N] [--steps K]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import bench
    import libprosic_b200 as P
    args = argparse.Namespace(sites=7168, reads_per_site=140, as_reads=a.reads, steps=a.steps)
    w = bench.gen_workload(args, 0)
    ctx = P.Context(0)
    dev = torch.device("cuda", 0)
    os.environ["VLR_TRACE"] = "1"
    res = bench.bench_allele_support(args, ctx, w, P.HMM_DEFAULT, 1, 1, torch, None, dev)
    print(res)


if __name__ == "__main__":
    main()
