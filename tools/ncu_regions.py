"""Split the stall samples / warp instructions of mpc_chain_kernel (an `ncu --page source --csv --print-source cuda,sass`
export) over the kernel's code regions: stage-in, ADMM loop (forward, corner + hull rows, backward), residuals, epilogue.
Region boundaries are found from the marker comments of mpc_chain.cu, so the tool follows the source.
usage: python tools/ncu_regions.py export.csv [path/to/mpc_chain.cu]"""
import csv
import os
import sys

src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "walking-controllers_b200", "csrc", "mpc_chain.cu")
lines = open(src).read().split("\n")
def find(marker):
    for n, l in enumerate(lines, 1):
        if marker in l:
            return n
    raise SystemExit("marker not found: " + marker)
marks = [("kernel entry / stage-in / problem build", find("__global__ void __launch_bounds__(2 * IPW) mpc_chain_kernel")),
         ("ADMM set-up (lambdas: residuals, infeasibility tests)", find("// ------------------------------------------------------------------ ADMM (osqp_solve)")),
         ("loop: forward sweep", find("// ---- forward: right-hand side")),
         ("loop: corner + hull rows", find("// ---- corner: dense 4x4")),
         ("loop: backward sweep", find("// ---- backward: L' x~")),
         ("loop: termination / adaptive rho", find("// ---- termination / adaptive rho")),
         ("epilogue", find("// ------------------------------------------------------------------ epilogue"))]
end = find("// instances per CTA for horizon N")
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; fname = ""; agg = {m[0]: [0, 0] for m in marks}; agg["inlined helpers / other files"] = [0, 0]; tot = itot = 0
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].rsplit("/", 1)[-1] if len(r) > 1 else ""; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5 or r[2] != "-":
        continue
    try:
        w = int(r[4]); n = int(r[7]); ln = int(r[0])
    except ValueError:
        continue
    key = "inlined helpers / other files"
    if "mpc_chain" in fname and marks[0][1] <= ln < end:
        for name, start in marks:
            if ln >= start:
                key = name
    agg[key][0] += w; agg[key][1] += n; tot += w; itot += n
print(f"total samples {tot}, warp instructions {itot}")
print("| region | samples | warp instructions |\n|---|---|---|")
for k, (w, n) in agg.items():
    print(f"| {k} | {100 * w / max(tot, 1):.1f} % | {100 * n / max(itot, 1):.1f} % |")
