"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` dump into the phases
between barriers (usage: python tools/ncu_phases.py source.csv [section]).  Diagnosis tooling."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
sec=int(sys.argv[2]) if len(sys.argv)>2 else 0
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
print([rows[i][1][:60] for i in starts])
a=starts[sec]; b=starts[sec+1] if sec+1<len(starts) else len(rows)
hdr=rows[a+1]; data=[r for r in rows[a+2:b] if len(r)==len(hdr)]
ix={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[ix['# Samples']]) for r in data)
print('total samples',tot, 'instr', len(data))
phase=[]; 
def new(k): return {'n':0,'samples':0,'exec':0,'st':{c:0 for c in stall_cols},'ops':{} ,'start':k}
cur=new(0)
for k,r in enumerate(data):
    toks=r[ix['Source']].split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    cur['n']+=1; cur['samples']+=int(r[ix['# Samples']]); cur['exec']+=int(r[ix['Instructions Executed']])
    for c in stall_cols: cur['st'][c]+=int(r[ix[c]] or 0)
    base=op.split('.')[0]
    cur['ops'][base]=cur['ops'].get(base,0)+int(r[ix['Instructions Executed']])
    if base=='BAR' or base=='EXIT':
        cur['end']=k; phase.append(cur); cur=new(k+1)
phase.append(cur)
for p in phase:
    if p['samples']<tot*0.005: continue
    top=sorted(p['st'].items(), key=lambda x:-x[1])[:6]
    ops=sorted(p['ops'].items(), key=lambda x:-x[1])[:9]
    print(f"instr {p['start']:5d}-{p.get('end',0):5d} n={p['n']:5d} samples={p['samples']:6d} ({100*p['samples']/tot:4.1f}%) exec={p['exec']/1e6:7.1f}M  ", ' '.join(f"{k[6:]}={v}" for k,v in top))
    print('      ', ' '.join(f"{k}={v/1e6:.1f}M" for k,v in ops))
