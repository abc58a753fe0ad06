#!/usr/bin/env python
"""bench.py — headline benchmark of the batched MPC + QP-IK controller tick on B200.

Metric (BASELINE.json): combined MPC+QP-IK controller ticks/sec at B=65536 iCub instances,
plus the batch tick latency distribution (p50/p99).  One step = one full control tick
(wc_mpc_tick + wc_ik_tick) of every instance of the batch on synthetic iCub walking inputs
(walking-controllers_b200/synthetic.py).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* `value`   device-resident inputs, CUDA-event timed, max over ranks (weak scaling: B per GPU fixed).
* `e2e`     same ticks through the host-pointer C ABI (wc_*_tick_host): pinned host inputs -> H2D,
            kernels, D2H of the results, inside the timed region.
* `roofline` dominant kernel (the ADMM kernel) against the measured HBM peak — reported as the
            contract asks, although the path is FP64-latency bound: the FP64 view is in
            `fp64`.
* `cpu_baseline` / `--impl reference`: the oracle restatement of the reference CPU path
            (oracle/, the reference itself cannot be built here), one instance per thread.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 65536
HORIZON = 25
P1_ITERS = 25         # phase 1 of the chain ADMM runs to the first termination check (check_termination = 25)
IK_RING = 8            # distinct IK input sets cycled over the ticks (IK is stateless per tick)
CPU_SAMPLE_B = 2048    # bounded CPU sample: instances x ticks
CPU_SAMPLE_TICKS = 12


def bytes_per_tick(N):
    """algorithmic HBM bytes per instance-tick"""
    return dict(mpc=128 * N + 464, ik=1540)


F_IK_FAST = 12500  # FK + Jacobian columns 6.0k, b/g build 0.5k, Woodbury + 9x9 Schur solve 6.0k


F_IK_SECOND = 6000  # one more Woodbury + Schur solve with a variable fixed at its bound (the FK is not counted twice)


def flops_per_tick(N, iters, refactors, wsc, slow_frac=1.0, second_frac=0.0):
    """algorithmic fp64 flops per instance-tick; the IK term is the mix of the fused-tick
    fast path (second_frac of the instances with its second pass: first working-set change in closed form) and the
    general null-space / active-set path (slow_frac of the instances; all of them when the fast path is off), whose
    working-set changes cost 3.9 kflop each"""
    n, m, nnzA = 4 * N + 2, 2 * N + 10, 6 * N + 18
    f_iter = 4 * nnzA + 21 * n + 12 * (n + m)
    wsc_general = max(wsc - max(second_frac - slow_frac, 0.0), 0.0)  # changes taken by the general kernel
    f_ik = (1.0 - slow_frac) * F_IK_FAST + second_frac * F_IK_SECOND + slow_frac * 43500 + 3900 * wsc_general
    return iters * f_iter + refactors * 36 * n, f_ik


METRIC = "MPC+QP-IK controller ticks/sec at B=65536 iCub instances"   # BASELINE.json metric (p99 latency: tick_latency_ms)


def workload_config(world, B, N):
    """`config` of both arms (ours and --impl reference), key for key: BASELINE configs[2]"""
    return {"workload": f"full tick: DCM-MPC (horizon {N}) + FK + QP-IK, iCub 23-DoF + floating base (BASELINE configs[2])",
            "instances_per_gpu": B, "instances_total": world * B, "horizon": N,
            "l2": "per-tick working set (solver state 10 KB x 65536 + inputs) exceeds the 126 MB L2; inputs change every tick"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, nm in enumerate(names):
                if len(r) > 5 + k and "Active" in r[5 + k] and "Not" not in r[5 + k]:
                    reasons.add(nm)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def ncu_counter(kernel, key="bytes_per_launch"):
    """per-launch counters of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json):
    DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and shared-memory wavefronts
    (l1tex__data_pipe_lsu_wavefronts_mem_shared.sum)"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get(key)
    except Exception:
        return None


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(q * (len(v) - 1))))]


# ------------------------------------------------------------------------------------------------
def run_reference(args, world, rank):
    """--impl reference: the reference's CPU path (oracle restatement, all host threads) on a bounded
    sample of the same workload.  Rank 0 only."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import oracle as O
    syn = importlib.import_module("walking-controllers_b200.synthetic")
    O.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    N = args.horizon
    steps, warm = args.steps, args.warmup
    T = steps + warm
    # instances per step: the workload's own batch when the whole run then fits ~3 minutes at the port's ~11 k ticks/s
    # per core, else a bounded sample of it (stated in `sample_instances_per_step`, outside `config`)
    budget = int(180.0 * 11000.0 * cores / max(T, 1))
    B = max(CPU_SAMPLE_B, min(args.batch, budget // 1024 * 1024))
    g = syn.GaitBatch(B, T, N)
    traj = g.trajectory().numpy()
    model = O.model_icub_like(); icfg = O.ik_default_config()
    qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = O.fk_batch(model, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    ik_in = [{k: v.numpy() for k, v in syn.ik_inputs(B, 100 + r, qreg, fkb, perturb=0.5).items()} for r in range(2)]
    mb = O.MpcBatch(O.mpc_default_config(N), B)
    feet = [g.feet(t) for t in range(T)]
    dcm = [g.measured_dcm(t).numpy() for t in range(T)]
    times = []
    for t in range(T):
        f = feet[t]
        a = (f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(), dcm[t],
             np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
        t0 = time.perf_counter()
        mb.tick(*a, threads=cores)
        O.ik_batch_tick(icfg, model, ik_in[t % 2], threads=cores)
        dt = time.perf_counter() - t0
        if t >= warm:
            times.append(dt)
    tot = sum(times)
    val = B * len(times) / tot
    sample = f"{B} instances x {len(times)} ticks of the same synthetic workload (N={N}), one instance per thread"
    cfg = workload_config(world, args.batch, N)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "ticks/s",
            "sample_instances_per_step": B,
            "note": ("CPU restatement of the reference path (osqp-eigen / qpOASES / iDynTree cannot be built offline), one "
                     f"instance per thread on {cores} host threads, {B} instances of the workload per step"),
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": tot / len(times) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "ticks/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(N):
    import numpy as np
    import torch
    from oracle import oracle as O
    syn = importlib.import_module("walking-controllers_b200.synthetic")
    O.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    B, T = CPU_SAMPLE_B, CPU_SAMPLE_TICKS
    g = syn.GaitBatch(B, T, N)
    traj = g.trajectory().numpy()
    model = O.model_icub_like(); icfg = O.ik_default_config()
    qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        o = O.fk_batch(model, ql.numpy(), bH.numpy(), lif.numpy(), jacobians=False)
        return {k: torch.from_numpy(v) for k, v in o.items()}

    ik_in = {k: v.numpy() for k, v in syn.ik_inputs(B, 100, qreg, fkb, perturb=0.5).items()}
    mb = O.MpcBatch(O.mpc_default_config(N), B)
    tot = 0.0
    t_mpc = t_ik = 0.0
    for t in range(T):
        f = g.feet(t)
        a = (f["lf_H"].numpy(), f["rf_H"].numpy(), f["l_contact"].numpy(), f["r_contact"].numpy(),
             g.measured_dcm(t).numpy(), np.ascontiguousarray(traj[2 * t: 2 * t + 2 * (N + 1)]))
        t0 = time.perf_counter()
        mb.tick(*a, threads=cores)
        t1 = time.perf_counter()
        O.ik_batch_tick(icfg, model, ik_in, threads=cores)
        t2 = time.perf_counter()
        if t >= 2:
            tot += t2 - t0; t_mpc += t1 - t0; t_ik += t2 - t1
    n = T - 2
    return {"value": B * n / tot, "unit": "ticks/s", "cores": cores, "kind": "port",
            "sample": f"{B} instances x {n} ticks (after 2 warm-up) of the same workload, one instance per thread; "
                      f"MPC {t_mpc / n * 1e3:.1f} ms + IK {t_ik / n * 1e3:.1f} ms per batch tick"}


# ------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(gpu_index):
    """Bind this rank's process (hence the first touch of its pinned host buffers) to the CPUs NVML reports as local to
    its GPU: with one rank per GPU the host->device copies of the e2e arm then read NUMA-local memory instead of crossing
    the socket interconnect.  Returns the number of CPUs kept, or None when the affinity could not be applied."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def build_closed_loop(wc, syn, torch, g, B, N, n_ticks, dev, local_rank, mpc_settings=None):
    """The closed-loop workload (synthetic.py CL_*): producer + MPC + glue + IK handles on the gait of `g`, the posture of
    syn.prepare_posture, a wc_loop handle over them.  Returns (loop, handles, configs)."""
    h_cl = syn.CL_COM_HEIGHT
    tcfg = wc.traj_default_config(n_ticks + 8, N); tcfg.com_height = h_cl
    mcfg_cl = wc.mpc_default_config(N); mcfg_cl.com_height = h_cl
    for k, v in (mpc_settings or {}).items():
        if k == "hull_tolerance":
            mcfg_cl.hull_tolerance = v
        else:
            setattr(mcfg_cl.osqp, k, v)
    gcfg = wc.glue_default_config(); gcfg.com_height = h_cl
    icfg_cl, icfg_open = wc.ik_default_config(), wc.ik_default_config()
    for j in range(23):
        icfg_cl.min_vel[j], icfg_cl.max_vel[j] = -syn.CL_VMAX, syn.CL_VMAX
        icfg_open.min_vel[j], icfg_open.max_vel[j] = -50.0, 50.0
    model = wc.model_icub_like()
    qreg = torch.tensor(list(icfg_cl.joint_reg_rad[:23]), dtype=torch.float64)
    tj = wc.TrajBatch(tcfg, B, device=local_rank)
    tj.set_gait(g.step_len, g.phase.to(torch.int32), g.dtheta, g.theta0, g.first_left.to(torch.uint8))
    ik_open = wc.IkBatch(icfg_open, model, B, device=local_rank)
    f0 = tj.feet(0); ref0, _ = tj.dcm(0, 1, want_vel=False)
    tgt = torch.cat([ref0[0:2], torch.full((1, B), h_cl, dtype=torch.float64, device=dev)], 0).contiguous()
    q0 = syn.prepare_posture(lambda q, b, l: ik_open.fk(q, b, l, jacobians=False), lambda inp: ik_open.tick(inp), qreg, f0, tgt)
    torch.cuda.synchronize()
    ik_open.close()
    mpc_cl = wc.MpcBatch(mcfg_cl, B, device=local_rank)
    ik_cl = wc.IkBatch(icfg_cl, model, B, device=local_rank)
    glue = wc.GlueBatch(gcfg, B, device=local_rank)
    loop = wc.LoopBatch(mpc_cl, ik_cl, glue, tj, device=local_rank)
    return loop, (mpc_cl, ik_cl, glue, tj), (mcfg_cl, icfg_cl, model), q0


def capi_collect(wc, torch, dist, world, rank, local_rank, B, zmp, dq, osqp_status, iters, qp_status, nwsr, stats_sum):
    """Result collection through the C ABI (wc_comm_* / wc_gather_*, include/wc_capi.h): all-gather of the per-instance results
    and all-reduce of the statistics vector over NCCL on the communicator's side stream.  torch.distributed only carries the
    128-byte NCCL id to the other ranks.  Returns (gathered dict, milliseconds of the timed collection)."""
    box = [wc.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = wc.Comm(box[0], world, rank, device=local_rank)
    # the first collective sets up the NCCL connections (seconds at 8 ranks): not part of the collection time
    w = torch.zeros(1, dtype=torch.float64, device=stats_sum.device)
    comm.gather(B, 0, stats_sum=w, wait="host")
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    out = comm.gather(B, dq.shape[0], zmp=zmp, dq=dq, osqp_status=osqp_status, iters=iters, qp_status=qp_status, nwsr=nwsr,
                      stats_sum=stats_sum, wait="host")
    ms = (time.perf_counter() - t0) * 1e3
    comm.close()
    return out, ms


MC_METRIC = "MPC+QP-IK controller ticks/sec, Monte-Carlo push / gain sweep at B=524288 iCub instances (closed loop)"


def run_mc_push(args, world, rank, local_rank):
    """--config mc_push: BASELINE configs[3].  524 288 instances in total, sharded over the ranks; every instance runs the
    closed-loop tick (wc_loop_*) with its own Q, R, k_posFoot, k_attFoot and one impulsive DCM push U(0, 0.06) m in a random
    direction at a random tick.  Results and statistics are collected with NCCL after the timed region."""
    import numpy as np
    import torch
    import torch.distributed as dist
    wc = importlib.import_module("walking-controllers_b200")
    syn = importlib.import_module("walking-controllers_b200.synthetic")
    wc.lib()
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = args.batch if args.batch != B_PER_GPU else 524288
    B, N = total // world, args.horizon
    steps, warm = args.steps, args.warmup
    T = warm + steps
    g = syn.GaitBatch(B, T, N, seed=0x5EED + rank, device=dev)
    loop, handles, (mcfg_cl, icfg_cl, model), q0 = build_closed_loop(wc, syn, torch, g, B, N, T, dev, local_rank)
    mpc_cl, ik_cl, glue, tj = handles
    rng = np.random.default_rng(0xC0FF + rank)
    q = 7500.0 * rng.choice([0.1, 0.3, 1, 3, 10], size=B); r = 9e6 * rng.choice([0.1, 1, 10], size=B)
    mpc_cl.set_weights(np.ascontiguousarray(np.stack([q, q])), np.ascontiguousarray(np.stack([r, r])))
    ik_cl.set_gains(np.ascontiguousarray(rng.choice([2.5, 4.0, 7.0], size=B)), np.ascontiguousarray(rng.choice([2.0, 5.0], size=B)))
    mag = rng.uniform(0, 0.06, B); ang = rng.uniform(0, 2 * np.pi, B)
    loop.set_push(torch.from_numpy(rng.integers(0, T, B).astype(np.int32)).to(dev),
                  torch.from_numpy(np.ascontiguousarray(np.stack([mag * np.cos(ang), mag * np.sin(ang)]))).to(dev))
    loop.reset(q0)
    view = loop.view()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # tails are tallied on the device, tick by tick, outside the graph (a handful of small torch kernels per tick)
    tally = torch.zeros(5, dtype=torch.float64, device=dev)  # not SOLVED, outside hull, IK infeasible, active set, iterations

    def tick_and_tally():
        loop.run(1)
        tally[0] += (view["osqp_status"] != 1).sum(); tally[1] += ((view["flags"] & 8) != 0).sum()
        tally[2] += (view["qp_status"] != 0).sum(); tally[3] += (view["nwsr"] > 0).sum(); tally[4] += view["iters"].sum()

    for t in range(warm):
        tick_and_tally()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(steps)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    tally[0] += (view["osqp_status"] != 1).sum(); tally[1] += ((view["flags"] & 8) != 0).sum()
    tally[2] += (view["qp_status"] != 0).sum(); tally[3] += (view["nwsr"] > 0).sum(); tally[4] += view["iters"].sum()
    gather_ms = None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        _, gather_ms = capi_collect(wc, torch, dist, world, rank, local_rank, B, view["zmp"], view["dq"], view["osqp_status"],
                                    view["iters"], view["qp_status"], view["nwsr"], tally)
    ms = ms.item()
    if rank == 0:
        n_tallied = float(world * B * (warm + 1))
        tl = tally.tolist()
        line = {"metric": MC_METRIC, "value": world * B * steps / (ms * 1e-3), "unit": "ticks/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "Monte-Carlo push-disturbance / gain sweep in closed loop (BASELINE configs[3])",
                           "instances_per_gpu": B, "instances_total": world * B, "horizon": N,
                           "sweep": "Q in 7500 x {0.1, 0.3, 1, 3, 10}, R in 9e6 x {0.1, 1, 10}, k_posFoot in {2.5, 4, 7}, "
                                    "k_attFoot in {2, 5}; push U(0, 0.06) m, random direction, random tick",
                           "l2": "solver state 10 KB x instances exceeds the 126 MB L2"},
                "tails": {"ticks_tallied_per_instance": warm + 1, "osqp_not_solved_frac": tl[0] / n_tallied,
                          "outside_hull_frac": tl[1] / n_tallied, "qpik_infeasible_frac": tl[2] / n_tallied,
                          "qpik_active_set_frac": tl[3] / n_tallied, "osqp_iters_mean": tl[4] / n_tallied,
                          "note": "tallied on the warm-up ticks and the last timed tick (the timed ticks are one graph replay each); "
                                  "parity of every tail against the oracle: tests/test_closed_loop.py"},
                "gpu_launches": int(steps), "api": "wc_loop_run_dev (one CUDA-graph replay per tick)", "clocks": clocks}
        if gather_ms is not None:
            line["nccl_gather_ms"] = gather_ms; line["nccl_gather_api"] = "wc_gather_async (C ABI, side stream)"
        print(json.dumps(line), flush=True)
    loop.close()
    for h in handles:
        h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours(args, world, rank, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    wc = importlib.import_module("walking-controllers_b200")
    syn = importlib.import_module("walking-controllers_b200.synthetic")
    wc.lib()  # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    numa_cpus = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = args.batch, args.horizon
    steps, warm = args.steps, args.warmup
    T = warm + steps

    # ---------------- synthetic inputs, resident in HBM before the timed region
    g = syn.GaitBatch(B, T, N, seed=0x5EED + rank, device=dev)
    traj = g.trajectory()
    mpc_in = []
    for t in range(T):
        f = g.feet(t)
        mpc_in.append((f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t)))
    mcfg = wc.mpc_default_config(N)
    icfg = wc.ik_default_config()
    model = wc.model_icub_like()
    ik = wc.IkBatch(icfg, model, B, device=local_rank)
    qreg = torch.tensor(list(icfg.joint_reg_rad[:23]), dtype=torch.float64)

    def fkb(ql, bH, lif):
        return ik.fk(ql, bH, lif, jacobians=False)

    ik_in = [syn.ik_inputs(B, 1000 * rank + r, qreg, fkb, device=dev, perturb=0.5) for r in range(IK_RING)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- arm 1: device-resident inputs
    mpc = wc.MpcBatch(mcfg, B, device=local_rank)
    mo = io_ = None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # before the warm-up: nvidia-smi needs ~100 ms for its first row, the timed region may be shorter
    for t in range(warm):
        a = mpc_in[t]
        mo = mpc.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=mo)
        io_ = ik.tick(ik_in[t % IK_RING], out=io_)
    barrier()
    mpc.enable_profiling(True); ik.enable_profiling(True)
    l0 = mpc.launches + ik.launches
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    it_sum = torch.zeros((), dtype=torch.float64, device=dev)
    it_max = torch.zeros((), dtype=torch.int32, device=dev)
    new_sum = torch.zeros((), dtype=torch.float64, device=dev)
    ws_sum = torch.zeros((), dtype=torch.float64, device=dev)
    evs[0].record()
    for k in range(steps):
        t = warm + k
        a = mpc_in[t]
        mo = mpc.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=mo)
        io_ = ik.tick(ik_in[t % IK_RING], out=io_)
        evs[k + 1].record()
    launches = mpc.launches + ik.launches - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]
    total_ms = evs[0].elapsed_time(evs[steps])
    prep_ms, admm_ms, nt = mpc.kernel_times_ms()
    setup_ms, phase1_ms, phase2_ms = mpc.kernel_times_detail_ms()
    ik_ms, nik = ik.kernel_times_ms()
    ik_fast_ms, ik_general_ms = ik.kernel_times_detail_ms()
    mpc.enable_profiling(False); ik.enable_profiling(False)
    # statistics of the last tick (outside the timed region)
    it_mean = mo["iters"].double().mean().item(); it_max = int(mo["iters"].max().item())
    solved = (mo["status"] == 1).double().mean().item()
    new_frac = ((mo["flags"] & 16) != 0).double().mean().item()
    ik_ok = (io_["status"] == 0).double().mean().item()
    wsc_mean = io_["nwsr"].double().mean().item()
    ws_frac = (io_["nwsr"] > 0).double().mean().item()
    ik_second, ik_general_n = ik.path_counts()   # last tick: second pass of the fast kernel / general kernel

    t_ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms_max = t_ms.item()

    # ---------------- NCCL result collection (off the tick's critical path)
    gather_ms = None
    if world > 1:
        stats = torch.tensor([solved, ik_ok, it_mean, wsc_mean], dtype=torch.float64, device=dev)
        _, gather_ms = capi_collect(wc, torch, dist, world, rank, local_rank, B, mo["zmp"], io_["dq"], mo["status"], mo["iters"],
                                    io_["status"], io_["nwsr"], stats)
        solved, ik_ok, it_mean, wsc_mean = (stats / world).tolist()

    # ---------------- arm 2: the closed-loop tick (wc_loop_*): producer -> measured CoM -> MPC -> ZMP-CoM law -> FK + QP-IK ->
    # integrator with the data dependence of WalkingModule::updateModule (the IK of tick t consumes the MPC output of tick t),
    # one CUDA-graph replay per tick, device resident.  Workload: synthetic.py CL_* (the model's own LIPM height, +-2 rad/s).
    h_cl = syn.CL_COM_HEIGHT
    # the pipeline's start (one upload with nothing beside it) and drain (the last tick's kernels and download) lie inside the timed
    # region: 40 steps keep them at ~3 % of it (20 at the larger batches and with several ranks on one box: pinned memory)
    e2e_steps = min(steps, 40 if (B <= 65536 and world == 1) else 20)   # (N > 1: host-link bound, and N x the pinned memory on one box)
    e2e_warm = min(warm, 5)
    Te = e2e_steps + e2e_warm
    cl_steps = min(steps, 200)
    loop, (mpc_cl, ik_cl, glue, tj), (mcfg_cl, icfg_cl, _), q0 = build_closed_loop(wc, syn, torch, g, B, N, Te + warm + cl_steps,
                                                                                    dev, local_rank)
    loop.reset(q0)
    view = loop.view()

    def pin(x):
        return x.detach().cpu().contiguous().pin_memory()

    # the first Te ticks are recorded: what a host that feeds the two blocks would upload, and what it would read back
    h_in, cl_out = [], []
    for t in range(Te):
        q_before = view["q_des"].clone()
        loop.run(1)
        torch.cuda.synchronize()
        # one pinned arena per tick in the packed order of wc_tick_host_async (wc_capi.h): two uploads per tick instead of sixteen
        rec = wc.host_tick_arena_in(B, N, 23, window=(t == 0))
        srcs = dict(lf_H=view["lf_H"], rf_H=view["rf_H"], l_contact=view["l_contact"], r_contact=view["r_contact"],
                    left_is_fixed=view["left_is_fixed"], dcm_meas=view["dcm_meas"], q_meas=q_before, q_lin=q_before,
                    lf_twist=view["lf_twist"], rf_twist=view["rf_twist"], com_pos_des=view["com_pos_des"],
                    com_vel_des=view["com_vel_des"], com_meas=view["com_meas"], neck_R_des=view["neck_R_des"])
        if t == 0:
            srcs["ref_window"] = view["ref_pos"]
        else:
            srcs["ref_newest"] = view["ref_pos"][2 * N: 2 * N + 2]
        for k_, v_ in srcs.items():
            rec[k_].copy_(v_.reshape(rec[k_].shape))
        h_in.append(rec)
        cl_out.append((view["zmp"].clone(), view["dq"].clone()))
    for t in range(warm):
        loop.run(1)
    barrier()
    ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ce0.record()
    loop.run(cl_steps)
    ce1.record()
    barrier()
    cl_ms = torch.tensor([ce0.elapsed_time(ce1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cl_ms, op=dist.ReduceOp.MAX)
    cl_ms = cl_ms.item()
    cl_stats = dict(osqp_solved_frac=(view["osqp_status"] == 1).double().mean().item(),
                    qpik_success_frac=(view["qp_status"] == 0).double().mean().item(),
                    qpik_active_set_frac=(view["nwsr"] > 0).double().mean().item())

    # ---------------- arm 3: end to end through the host-pointer C ABI (wc_tick_host_async, pinned host buffers): the
    # recorded closed-loop ticks replayed by a host that uploads every tick's inputs and reads every tick's results back
    mpc2 = wc.MpcBatch(mcfg_cl, B, device=local_rank)
    ik2 = wc.IkBatch(icfg_cl, model, B, device=local_rank)

    DEPTH = 3
    h_outs = [wc.host_tick_arena_out(B, 23) for _ in range(DEPTH)]   # ticks in flight -> as many result arenas (one download each)
    h2d = sum(v.numel() * v.element_size() for k, v in h_in[1].items() if k != "_arena")
    d2h = sum(v.numel() * v.element_size() for k, v in h_outs[0].items() if k != "_arena")
    e2e_match = True

    def submit(t):
        # one packed upload for both blocks (shared feet transforms, stance base selected on the device, one DCM sample per
        # tick), MPC and IK kernels on the handle's two kernel streams, one download.  Staging is double-buffered: the upload
        # of tick t+1 and the download of tick t-1 overlap the kernels of tick t.
        mpc2.tick_fused_host_async(ik2, h_in[t], h_outs[t % DEPTH])

    def retire(t):
        nonlocal e2e_match
        mpc2.sync_one()   # results of tick t are now in its pinned host buffers
        if t < e2e_warm:  # (outside the timed region) the replay reproduces the closed loop bit for bit
            zc, dc = cl_out[t]
            e2e_match = e2e_match and torch.equal(h_outs[t % DEPTH]["zmp"], zc.cpu()) and torch.equal(h_outs[t % DEPTH]["dq"], dc.cpu())

    for t in range(e2e_warm):
        submit(t)
        if t >= DEPTH - 1:
            retire(t - (DEPTH - 1))
    mpc2.sync()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        submit(e2e_warm + k)
        if k >= DEPTH - 1:
            mpc2.sync_one()
    mpc2.sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    zc, dc = cl_out[Te - 1]
    e2e_match = e2e_match and torch.equal(h_outs[(Te - 1) % DEPTH]["zmp"], zc.cpu()) and torch.equal(h_outs[(Te - 1) % DEPTH]["dq"], dc.cpu())
    e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    e2e_s = e.item()
    barrier()

    # ---------------- SURVEY §8(f) rows built next to the headline path: device times of one tick at the same batch
    # (outside the timed regions above; N = 1 only)
    next_rows = None
    if world == 1 and not args.no_extras:
        def timed(fn, n_warm=3, n=10):
            for k in range(n_warm):
                fn(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(n):
                fn(n_warm + k)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        ik2 = wc.IkBatch(icfg, model, B, device=local_rank)
        ik2.set_solver(wc.SOLVER_OSQP)
        o2 = {}
        def f_osqp(k):
            o2["o"] = ik2.tick(ik_in[k % IK_RING], out=o2.get("o"))
        osqp_ms = timed(f_osqp)
        osqp_it = o2["o"]["nwsr"].double().mean().item(); osqp_ok = (o2["o"]["status"] == 1).double().mean().item()
        ik2.close()
        tj2 = wc.TrajBatch(wc.traj_default_config(T, N), B, device=local_rank)
        tj2.set_gait(g.step_len, g.phase.to(torch.int32), g.dtheta, g.theta0, g.first_left.to(torch.uint8))
        tb = {}
        def f_traj(k):
            tb["f"] = tj2.feet(k, out=tb.get("f"))
            tb["p"], _ = tj2.dcm(k, N + 1, pos=tb.get("p"), want_vel=False)
            tb["m"] = tj2.measured_dcm(k, 0.002, 0x5EED, out=tb.get("m"))
        traj_ms = timed(f_traj)
        tj2.close()
        gl = wc.GlueBatch(wc.glue_default_config(), B, device=local_rank)
        z2 = torch.zeros(2, B, dtype=torch.float64, device=dev); c3 = torch.zeros(3, B, dtype=torch.float64, device=dev)
        gb = {}
        def f_glue(k):
            gb["c"] = gl.com_reference(z2, mo["zmp"], z2, c3, out=gb.get("c"))
            gb["q"] = gl.integrate_joints(io_["dq"], out=gb.get("q"))
            gb["r"] = gl.reactive_dcm(1.5, z2, z2, z2, out=gb.get("r"))
        glue_ms = timed(f_glue)
        gl.close()
        next_rows = {"ik_osqp_tick_ms": osqp_ms, "ik_osqp_iters_mean": osqp_it, "ik_osqp_solved_frac": osqp_ok,
                     "ik_osqp_ticks_per_s": B / (osqp_ms * 1e-3),
                     "traj_producer_tick_ms": traj_ms, "glue_tick_ms": glue_ms,
                     "note": "FK + WalkingQPIK_osqp tick (rank 2), on-device input producer: feet + DCM window + measured DCM "
                             "(rank 3), LIPM / ZMP-CoM law + joint integrator + reactive DCM law (ranks 1, 4); B instances, device "
                             "resident, CUDA events"}

    if rank == 0:
        hbm_peak, peak_src = peaks()
        fp64_peak = wc.measure_fp64_peak(local_rank)
        bt = bytes_per_tick(N)
        admm_avg_ms = admm_ms / max(nt, 1)
        ik_avg_ms = ik_ms / max(nik, 1)
        prep_avg_ms = prep_ms / max(nt, 1)
        # dominant kernel: the ADMM pass every instance goes through (chain path: phase 1; cooperative path: the
        # single ADMM kernel).  Its algorithmic bytes are the MPC share of SURVEY §8(d): (128 N + 464) B per instance.
        dom_name = {"split": "mpc_split_kernel<phase 1>", "chain": "mpc_chain_kernel<phase 1>"}.get(mpc.path, "mpc_admm_kernel")
        dom_ms = phase1_ms / max(nt, 1)
        achieved = bt["mpc"] * B / (dom_ms * 1e-3) * 1e-9
        at_capture_size = B == B_PER_GPU and N == HORIZON
        # shared-memory pipe of the dominant kernel: wavefronts counted by ncu (l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of
        # the committed capture, one wavefront = one 128-byte pass of the crossbar) over the kernel time measured HERE
        wf = ncu_counter(dom_name, "smem_wavefronts_per_launch") if at_capture_size else None
        smem_peak = 128 * 148 * (clocks["sm_max_mhz"] or 1965.0) * 1e-3 if clocks else 128 * 148 * 1.965
        smem_ach = wf * 128 / (dom_ms * 1e-3) * 1e-9 if wf else None
        # FP64 roofline of the same kernel with the flop model of SURVEY 8(d): iterations x F_iter(N) per instance
        f_p1 = flops_per_tick(N, float(P1_ITERS if mpc.path != "coop" else it_mean), 0.0, 0.0)[0]
        fast_on = ik_fast_ms > 0 and ik_general_ms > 0
        f_mpc, f_ik = flops_per_tick(N, it_mean, 0.0, wsc_mean, slow_frac=(ik_general_n / B) if fast_on else 1.0,
                                     second_frac=(ik_second / B) if fast_on else 0.0)
        value = world * B * steps / (total_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "ticks/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": total_ms_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, B, N),
            "mpc_path": mpc.path, **({"rank_cpu_affinity": numa_cpus} if world > 1 else {}),
            "tick_latency_ms": {"p50": pct(step_ms, 0.5), "p99": pct(step_ms, 0.99), "max": max(step_ms)},
            "kernels_ms": {"mpc_prepare": prep_avg_ms, "mpc_admm": admm_avg_ms, "fk_qpik": ik_avg_ms,
                           "mpc_setup": setup_ms / max(nt, 1), "mpc_phase1": phase1_ms / max(nt, 1),
                           "mpc_phase2": phase2_ms / max(nt, 1), "ik_fast": ik_fast_ms / max(nik, 1),
                           "ik_general": ik_general_ms / max(nik, 1)},
            "stats": {"osqp_solved_frac": solved, "osqp_iters_mean": it_mean, "osqp_iters_max_last_tick": it_max,
                      "new_solver_frac": new_frac, "qpik_success_frac": ik_ok, "qpik_nwsr_mean": wsc_mean,
                      "qpik_active_set_frac": ws_frac,
                      "qpik_second_pass_frac": ik_second / B, "qpik_general_kernel_frac": ik_general_n / B},
            "roofline": {"bound": "hbm", "kernel": dom_name, "kernel_ms": dom_ms, "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_counter(dom_name) if at_capture_size else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bt["mpc"] * B,
                         "note": "the path is not HBM bound (DESIGN.md §5): ~5 KB of compulsory traffic against ~0.12 Mflop of "
                                 "fp64 in dependent chains per instance-tick; the dominant kernel is bound by the shared-memory "
                                 "data pipe (the `smem` object; ncu: 72 % of its sustained peak), see also roofline_fp64.  traffic "
                                 "exceeds the algorithmic bytes by the Ruiz scaling vectors (D, 1/D, E, 1/E), read every tick"},
            "roofline_fp64": {"bound": "fp64", "kernel": dom_name, "kernel_ms": dom_ms,
                              "achieved": f_p1 * B / (dom_ms * 1e-3) * 1e-12, "peak": fp64_peak * 1e-3, "unit": "TFLOP/s",
                              "frac": f_p1 * B / (dom_ms * 1e-3) * 1e-9 / fp64_peak,
                              "flops_per_launch": f_p1 * B,
                              "peak_source": "measured here: wc_measure_fp64_peak, register-resident FMA chains on all SMs "
                                             "(MEASURED_PEAKS.json has no FP64 entry)",
                              "model": "SURVEY 8(d): iterations x (4 nnz(A) + 21 n + 12 (n + m)) flops per instance"},
            "smem": {"bound": "shared memory", "kernel": dom_name, "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s",
                     "frac": smem_ach / smem_peak if smem_ach else None, "wavefronts_per_launch": wf,
                     "peak_source": "128 B/clk/SM x 148 SMs x SM clock (B300_MICROARCH.md smem crossbar); wavefronts from the "
                                    "committed ncu capture (profiles/ncu_traffic.json)"},
            "fp64": {"peak_gflops_measured": fp64_peak,
                     "mpc_admm_gflops": f_mpc * B / (admm_avg_ms * 1e-3) * 1e-9,
                     "fk_qpik_gflops": f_ik * B / (ik_avg_ms * 1e-3) * 1e-9,
                     "mpc_admm_frac": f_mpc * B / (admm_avg_ms * 1e-3) * 1e-9 / fp64_peak,
                     "fk_qpik_frac": f_ik * B / (ik_avg_ms * 1e-3) * 1e-9 / fp64_peak,
                     "flops_per_tick_model": {"mpc": f_mpc, "ik": f_ik}},
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "ticks/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": "wc_tick_host_async (fused MPC + FK/QP-IK host tick, one pinned arena per direction, up to three ticks in flight)",
                    "workload": "the inputs of a recorded closed-loop rollout (wc_loop_*), uploaded tick by tick",
                    "matches_closed_loop_bitwise": bool(e2e_match)},
            # the closed-loop tick with the true MPC -> ZMP-CoM law -> IK dependence, one CUDA-graph replay per tick
            "closed_loop": {"value": world * B * cl_steps / (cl_ms * 1e-3), "unit": "ticks/s", "ms_per_tick": cl_ms / cl_steps,
                            "steps": cl_steps, "api": "wc_loop_run_dev (CUDA graph)", "stats": cl_stats,
                            "workload": f"periodic walking from the prepared posture, LIPM height {h_cl} m, |dq| <= {syn.CL_VMAX} rad/s"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if gather_ms is not None:
            line["nccl_gather_ms"] = gather_ms; line["nccl_gather_api"] = "wc_gather_async (C ABI, side stream)"
        if next_rows is not None:
            line["next_rows"] = next_rows
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg(N)
        print(json.dumps(line), flush=True)
    loop.close(); mpc.close(); mpc2.close(); ik.close(); ik2.close(); mpc_cl.close(); ik_cl.close(); glue.close(); tj.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the timings of the SURVEY 8(f) rows (next_rows)")
    ap.add_argument("--horizon", type=int, default=HORIZON, help="MPC horizon N (default: BASELINE configs[2], 25)")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="instances per GPU (default 65536; mc_push: in total, default 524288)")
    ap.add_argument("--config", default="full_tick", choices=["full_tick", "mc_push"],
                    help="full_tick = BASELINE configs[2] (the headline); mc_push = configs[3], Monte-Carlo push / gain sweep in closed loop")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if args.config == "mc_push":
        run_mc_push(args, world, rank, local_rank)
        return
    run_ours(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
