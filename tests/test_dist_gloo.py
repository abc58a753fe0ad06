"""World-size-2 `gloo` test of the multi-GPU plumbing on CPU: instance sharding, result gather and
statistics all-reduce (walking-controllers_b200/dist.py), with the oracle standing in for the device
tick (the data path has no collective: SURVEY.md §8(e))."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B_total, N, T, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    syn = importlib.import_module("walking-controllers_b200.synthetic")
    D = importlib.import_module("walking-controllers_b200.dist")
    lo, hi = D.shard_range(B_total, rank, world)
    g = syn.GaitBatch(B_total, T, N, seed=31)            # every rank derives the same global inputs ...
    traj = g.trajectory().numpy()
    mb = O.MpcBatch(O.mpc_default_config(N), hi - lo)     # ... and ticks only its shard
    for t in range(T):
        f = g.feet(t)
        sl = lambda a: np.ascontiguousarray(a.numpy()[..., lo:hi])
        z, st, it, fl = mb.tick(sl(f["lf_H"]), sl(f["rf_H"]), sl(f["l_contact"]), sl(f["r_contact"]),
                                sl(g.measured_dcm(t)), np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1), lo:hi]))
    res = D.gather_results(torch.from_numpy(z), world, rank)
    stats = D.reduce_stats(torch.from_numpy(st), torch.from_numpy(it), torch.zeros(hi - lo, dtype=torch.int32),
                           torch.zeros(hi - lo, dtype=torch.int32), world)
    if rank == 0:
        ret["zmp"] = res.numpy(); ret["stats"] = stats
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ticks_match_single_process(oracle, syn):
    B, N, T, world = 11, 10, 5, 2      # odd batch: shards of 5 and 6 instances
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, N, T, ret), nprocs=world, join=True)
    g = syn.GaitBatch(B, T, N, seed=31)
    traj = g.trajectory().numpy()
    mb = oracle.MpcBatch(oracle.mpc_default_config(N), B)
    for t in range(T):
        f = g.feet(t)
        z, st, it, fl = mb.tick(f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(),
                                g.measured_dcm(t).numpy(), np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
    assert np.array_equal(ret["zmp"], z)
    s = ret["stats"]
    assert s["n"] == B and s["mpc_solved"] == (st == 1).sum() and s["iters_sum"] == it.sum() and s["iters_max"] == it.max()


def test_shard_ranges_partition_the_batch():
    D = importlib.import_module("walking-controllers_b200.dist")
    for B in (1, 7, 64, 65536, 524288):
        for world in (1, 2, 4, 8):
            r = [D.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
