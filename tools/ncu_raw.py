"""Key `--set full` metrics of every launch in an .ncu-rep (raw page): python tools/ncu_raw.py file.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
for r in rows[2:]:
    print("###", r[hdr.index("Kernel Name")])
    for w in want:
        if w in hdr:
            print(f"- {w}: {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try: st.append((h.split("stalled_")[1].split("_per")[0], float(r[i].replace(",", ""))))
            except ValueError: pass
    st.sort(key=lambda x: -x[1])
    print("- top stalls (warps stalled per issue-active cycle): " + ", ".join(f"{k}={v:.2f}" for k, v in st[:8]))
    print()
