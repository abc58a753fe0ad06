"""Stall samples and executed warp instructions of mpc_split_kernel<0, 25> by code region (stage-in / ADMM loop / after the loop), from an ncu report with sources; the loop is found from the executed counts.  usage: python tools/ncu_split_regions.py report.ncu-rep"""
import csv, collections, re, sys, subprocess
rep=sys.argv[1]; 
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','--kernel-name','regex:mpc_split','--launch-skip','0','--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.split('\n')))
hdr=rows[1]
iS=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
st=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[]; seen=set()
for r in rows[2:]:
    try: ad=int(r[0],16)
    except: continue
    if ad in seen: continue
    seen.add(ad); data.append((ad,r))
base=data[0][0]
print(rows[0][1][:70], 'instr rows', len(data))
# find hot loop = rows executed >= 20*8192
hot=[ad-base for ad,r in data if int(r[iE])>=8192*20]
lo,hi=min(hot),max(hot)
print('hot loop', hex(lo),hex(hi),(hi-lo)//16+1)
tot=sum(int(r[iS]) for ad,r in data)
regions=[('prologue',0,lo),('loop',lo,hi+16),('after loop',hi+16,1<<30)]
for name,a,b in regions:
    c=collections.Counter(); n=0; ie=0
    for ad,r in data:
        if a<=ad-base<b:
            n+=int(r[iS]); ie+=int(r[iE])
            for i,h in st: c[h]+=int(r[i] or 0)
    print(f'{name:12s} samples {n:6d} ({100*n/tot:.1f}%) instr/warp {ie/8192:.0f}', [(k[6:],v) for k,v in c.most_common(6)])
# after-loop buckets
b=collections.OrderedDict()
for ad,r in data:
    o=ad-base
    if o>hi:
        k=(o)//(16*256)
        s=b.setdefault(k,[0,0]); s[0]+=int(r[iS]); s[1]+=int(r[iE])
for k,v in b.items(): print(hex(k*16*256), v[0], round(v[1]/8192))
