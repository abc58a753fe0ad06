"""Multi-GPU plumbing (torch.distributed): the batch of independent controller instances shards
trivially across ranks; there is no collective inside a tick.  Collectives are used
only to gather per-instance results on rank 0 and to all-reduce a small statistics vector."""
import torch
import torch.distributed as dist


def shard_range(B_total, rank, world):
    """contiguous instance range [lo, hi) owned by `rank`"""
    lo = B_total * rank // world
    hi = B_total * (rank + 1) // world
    return lo, hi


def gather_results(local, world, rank, dst=0):
    """local: [rows, B_local] tensor (field-major).  Returns [rows, B_total] on dst, None elsewhere.
    Shards may differ in size by one instance: sizes are exchanged first."""
    if world == 1:
        return local
    n = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = torch.zeros(local.shape[0], nmax, dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:, :s] for b, s in zip(bufs, sizes)], dim=1)


def reduce_stats(status_mpc, iters, status_ik, nwsr, world):
    """all-reduced statistics vector: [n, n_solved, sum_iters, max_iters, n_ik_ok, sum_nwsr, max_nwsr]"""
    dev = iters.device
    s = torch.stack([torch.tensor(float(iters.numel()), device=dev, dtype=torch.float64),
                     (status_mpc == 1).double().sum(), iters.double().sum(),
                     (status_ik == 0).double().sum(), nwsr.double().sum()])
    m = torch.stack([iters.double().max(), nwsr.double().max()])
    if world > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return dict(n=s[0].item(), mpc_solved=s[1].item(), iters_sum=s[2].item(), iters_max=m[0].item(),
                ik_ok=s[3].item(), nwsr_sum=s[4].item(), nwsr_max=m[1].item())
