"""N>1 host logic on CPU: region sharding + gather in record order, world size 2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_balance_and_groups():
    from libprosic_b200.sharding import shard_bounds
    rng = np.random.default_rng(0)
    w = rng.integers(1, 100, 1000).astype(float)
    for world in (1, 2, 4, 8):
        b = shard_bounds(w, world)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        loads = [w[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) <= w.sum() / world + 100
    # breakend groups never straddle a boundary
    g = -np.ones(1000, dtype=int)
    g[495:510] = 7
    b = shard_bounds(np.ones(1000), 2, g)
    assert not (495 < b[1] < 510)
    assert shard_bounds([], 4).tolist() == [0, 0, 0, 0, 0]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle as O
    from libprosic_b200 import scenario, synth
    from libprosic_b200.sharding import gather_in_record_order, my_slice, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank sees the same candidate list; each computes only its region slice (here with the
    # CPU oracle standing in for the per-rank GPU call) and rank 0 gathers in record order
    scn = scenario.single_sample(continuous=False)
    cols, cat, off, _ = synth.make_pileups(40, [12], seed=3, ragged=True)
    bounds = shard_bounds(np.diff(off) + 1, world)
    lo, hi = my_slice(bounds, rank)
    spec = O.ModelSpec(scn["n_samples"], scn["resolution"], scn["contaminated"], scn["by"],
                       scn["purity"], scn["nodes"], scn["events"],
                       [O.make_biases(**b) for b in scn["biases"]])
    local = np.array([O.model_compute(spec, [O.PileupView(cols, cat, int(off[s]), int(off[s + 1]))])["event_ln"]
                      for s in range(lo, hi)]).reshape(hi - lo, len(scn["events"]))
    full = gather_in_record_order(local, bounds, rank, world, dist)
    if rank == 0:
        want = np.array([O.model_compute(spec, [O.PileupView(cols, cat, int(off[s]), int(off[s + 1]))])["event_ln"]
                         for s in range(40)])
        q.put(bool(np.array_equal(full, want, equal_nan=True)) and full.shape == want.shape)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_in_record_order_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_bind_host_to_device_is_harmless_without_a_gpu():
    """No NVML / no device: the helper reports None and leaves the affinity mask alone."""
    import os
    from libprosic_b200.sharding import bind_host_to_device
    before = os.sched_getaffinity(0)
    got = bind_host_to_device(0)
    if got is None:
        assert os.sched_getaffinity(0) == before
    else:  # a GPU box: the mask shrank to the device's cores; restore it
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)
