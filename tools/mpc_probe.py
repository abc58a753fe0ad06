"""Quick throughput probe of the batched MPC (device-resident inputs). Not the bench."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
wc = importlib.import_module("walking-controllers_b200")
syn = importlib.import_module("walking-controllers_b200.synthetic")

def run(B, N, T=70, warm=20):
    dev = torch.device("cuda:0")
    g = syn.GaitBatch(B, T, N, device=dev)
    traj = g.trajectory()
    ins = []
    for t in range(T):
        f = g.feet(t)
        ins.append((f["lf_H"], f["rf_H"], f["l_contact"], f["r_contact"], g.measured_dcm(t)))
    m = wc.MpcBatch(wc.mpc_default_config(N), B)
    out = None
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
    iters = []
    for t in range(T):
        evs[t].record()
        a = ins[t]
        out = m.tick(a[0], a[1], a[2], a[3], a[4], traj[2 * t: 2 * t + 2 * (N + 1)], out=out)
        if t >= warm: iters.append(out["iters"].clone())
    evs[T].record()
    torch.cuda.synchronize()
    ms = [evs[t].elapsed_time(evs[t + 1]) for t in range(warm, T)]
    it = torch.stack(iters).double()
    st = out["status"]
    print(f"lanes={os.environ.get('WC_MPC_LANES','auto')} B={B} N={N}: mean {sum(ms)/len(ms):.3f} ms/tick  min {min(ms):.3f} max {max(ms):.3f}"
          f"  -> {B/(sum(ms)/len(ms))*1e3:.3e} ticks/s ; mean iters {it.mean().item():.1f} max {it.max().item():.0f}; solved {(st==1).float().mean().item():.4f}", flush=True)
    m.close()

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    run(B, N)
