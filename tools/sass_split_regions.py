"""Static size of the phase-1 split-chain kernel by region (before / in / after the ADMM loop) and the opcode mix of the loop, from cuobjdump.  usage: python tools/sass_split_regions.py [object or library]"""
import re,sys,subprocess,collections
obj=sys.argv[1] if len(sys.argv)>1 else 'walking-controllers_b200/csrc/build/mpc_split.o'
txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
parts=txt.split('Function : ')
for p in parts[1:]:
    name=p.split('\n')[0]
    if 'ILi0ELi25' not in name: continue
    ins=[]
    for l in p.split('\n'):
        m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2)))
    print(name, len(ins),'instr')
    loops=[]
    for a,t in ins:
        m=re.search(r'BRA(?:\.\w+)* .*?0x([0-9a-f]+)',t)
        if m:
            tgt=int(m.group(1),16)
            if tgt<a: loops.append((tgt,a)); print('  back',hex(a),'->',hex(tgt),(a-tgt)//16+1,'instr')
    def hist(lo,hi):
        c=collections.Counter()
        for a,t in ins:
            if lo<=a<=hi:
                t=re.sub(r'^@!?U?P\d+\s+','',t); c[t.split()[0].split('.')[0]]+=1
        return c
    # the hot loop = smallest loop with > 500 instr
    hot=[l for l in loops if 500<(l[1]-l[0])//16<2000]
    if hot:
        lo,hi=hot[0]; print('  hot loop',hex(lo),hex(hi),(hi-lo)//16+1, hist(lo,hi).most_common(10))
        print('  before loop',lo//16,'instr; after loop',(ins[-1][0]-hi)//16)
        print('  after-loop hist',hist(hi+16,ins[-1][0]).most_common(12))
