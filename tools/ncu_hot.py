"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: hottest source lines by stall samples."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None; data = {}; tot = 0; inst_tot = 0
fname = ''
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].rsplit('/', 1)[-1] if len(r) > 1 else ''; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[2] != '-':  # SASS rows
        continue
    try:
        w = int(r[4]); n = int(r[7])
    except Exception:
        continue
    ln = (fname + ':' if fname else '') + r[0]
    stall = {h: int(v) for h, v in zip(hdr[32:49], r[32:49]) if v.isdigit()}
    d = data.setdefault(ln, [0, 0, r[1][:100], {}])
    d[0] += w; d[1] += n
    for k, v in stall.items(): d[3][k] = d[3].get(k, 0) + v
    tot += w; inst_tot += n
print('total samples', tot, 'warp instructions', inst_tot)
agg = {}
for ln, (w, n, s, st) in data.items():
    for k, v in st.items(): agg[k] = agg.get(k, 0) + v
print('stall mix:', ', '.join(f"{k[6:]}={100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for ln, (w, n, s, st) in sorted(data.items(), key=lambda x: -x[1][0])[:top]:
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{w:7d} {100*w/max(tot,1):5.1f}% inst={n:10d} {ln:>22}: {s[:90]}  [{', '.join(k[6:]+':'+str(v) for k,v in main)}]")
