#!/usr/bin/env python
"""GCUPS of the long-read strip kernel (config C4 shape) on a small subsample, for tuning.

    python tools/c4_probe.py [--pairs N] [--steps K]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1184)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import bench
    import libprosic_b200 as P
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = P.Context(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def timed(fn, steps):
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)
    args = argparse.Namespace(c4_pairs=a.pairs, steps=a.steps, warmup=1, exact_sum=False)
    print(bench.bench_c4(args, ctx, torch, None, dev, 0, 1, timed))


if __name__ == "__main__":
    main()
