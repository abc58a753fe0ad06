"""Summarise a gpurun visit into profiles/: launch list shares (ncu --metrics gpu__time_duration.sum) and the key
`--set full` metrics of the hot kernels.  Usage: python tools/ncu_summary.py TAG  (reads gpurun_out/*_TAG.*)"""
import csv, json, os, subprocess, sys, collections
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out"); po = os.path.join(root, "profiles"); os.makedirs(po, exist_ok=True)
out = []
def short(n):
    for k in ("mpc_admm", "mpc_chain", "mpc_prepare", "mpc_setup", "ik_kernel", "ik_", "peak_fp64", "mpc_"):
        if k in n:
            return n.split("(")[0].replace("void ", "").replace("wcb::", "").replace("<unnamed>::", "")
    return "torch/other: " + n[:50]
lp = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        a = agg.setdefault(short(r[ki]), [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append(f"## launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 3 --warmup 12 --no-cpu`\n")
    out.append("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|")
    for k, (n, ns) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {n} | {ns/1e6:.3f} | {ns/n/1e3:.1f} | {100*ns/tot:.1f}% |")
    out.append("")
for rp in (os.path.join(go, f"prof_{tag}.ncu-rep"), os.path.join(go, f"prof_osqp_{tag}.ncu-rep")):
  if os.path.exists(rp):
      raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
      rows = list(csv.reader(raw.splitlines()))
      hdr, units = rows[0], rows[1]
      want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
              "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "lts__t_sector_hit_rate.pct"]
      out.append(f"## ncu --set full ({tag}), per captured launch\n")
      for r in rows[2:]:
          out.append(f"### {short(r[hdr.index('Kernel Name')])}\n")
          for w in want:
              if w in hdr:
                  out.append(f"- {w}: {r[hdr.index(w)]} {units[hdr.index(w)]}")
          st = [(h, float(r[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
          st.sort(key=lambda x: -x[1])
          out.append("- top stalls (warps stalled per issue-active cycle): " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for h, v in st[:7]))
          out.append("")
bp = os.path.join(go, f"bench_{tag}.json")
if os.path.exists(bp):
    out.append(f"## bench line of the same visit ({tag}, not under ncu)\n\n```\n" + open(bp).read().strip() + "\n```\n")
bp = os.path.join(go, f"bench_524k_{tag}.json")
if os.path.exists(bp):
    out.append(f"## 524 288 instances on one GPU (BASELINE configs[3] size; {tag})\n\n```\n" + open(bp).read().strip() + "\n```\n")
bp = os.path.join(go, f"bench_ref_{tag}.json")
if os.path.exists(bp):
    out.append(f"## --impl reference line ({tag})\n\n```\n" + open(bp).read().strip() + "\n```\n")
open(os.path.join(po, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])
