"""wc_comm_* / wc_gather_* (SURVEY 8(b), 8(e)): NCCL all-gather of the per-instance results and all-reduce of the statistics
vector through the C ABI, on the communicator's side stream.  One rank on one GPU always; two ranks when two GPUs are visible."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _payload(rank, B, nJ, dev):
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    zmp = torch.randn(2, B, generator=g, dtype=torch.float64).to(dev)
    dq = torch.randn(nJ, B, generator=g, dtype=torch.float64).to(dev)
    st = torch.randint(-4, 3, (B,), generator=g, dtype=torch.int32).to(dev)
    it = torch.randint(25, 200, (B,), generator=g, dtype=torch.int32).to(dev)
    return zmp, dq, st, it


def _rank_main(rank, world, path, B, nJ, q):
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    wc = importlib.import_module("walking-controllers_b200")
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    if rank == 0:
        with open(path + ".tmp", "wb") as f:
            f.write(wc.comm_unique_id())
        os.replace(path + ".tmp", path)
    else:
        import time
        while not os.path.exists(path):
            time.sleep(0.05)
    uid = open(path, "rb").read()
    c = wc.Comm(uid, world, rank, device=rank)
    zmp, dq, st, it = _payload(rank, B, nJ, dev)
    ssum = torch.tensor([float((st == 1).sum()), float(it.sum())], dtype=torch.float64, device=dev)
    smax = torch.tensor([float(it.max())], dtype=torch.float64, device=dev)
    out = c.gather(B, nJ, zmp=zmp, dq=dq, osqp_status=st, iters=it, stats_sum=ssum, stats_max=smax, wait="host")
    ok = True
    tot_solved = tot_it = 0.0; mx = 0.0
    for r in range(world):
        z, d, s, i = _payload(r, B, nJ, dev)
        ok = ok and torch.equal(out["zmp"][r], z) and torch.equal(out["dq"][r], d) and torch.equal(out["osqp_status"][r], s)
        ok = ok and torch.equal(out["iters"][r], i)
        tot_solved += float((s == 1).sum()); tot_it += float(i.sum()); mx = max(mx, float(i.max()))
    ok = ok and ssum.tolist() == [tot_solved, tot_it] and smax.tolist() == [mx]
    c.close()
    if q is not None:
        q.put((rank, bool(ok)))
    return ok


def test_gather_single_rank(wc, cuda_device):
    with tempfile.TemporaryDirectory() as d:
        assert _rank_main(0, 1, os.path.join(d, "id"), 1000, 23, None)


def test_gather_two_ranks(wc, cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        ps = [ctx.Process(target=_rank_main, args=(r, 2, os.path.join(d, "id"), 4096, 23, q)) for r in range(2)]
        for p in ps:
            p.start()
        res = [q.get(timeout=120) for _ in ps]
        for p in ps:
            p.join(timeout=60)
    assert all(ok for _, ok in res), res
