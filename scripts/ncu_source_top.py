import csv,sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdrs=[i for i,r in enumerate(rows) if r and r[0]=='Address']
hdr=rows[hdrs[0]]; idx={n:i for i,n in enumerate(hdr)}
stall=[n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
tot=Counter(); inst=[]
end = hdrs[1] if len(hdrs)>1 else len(rows)
for r in rows[hdrs[0]+1:end]:
    if len(r)<len(hdr): continue
    try: n=int(r[idx['# Samples']] or 0)
    except: continue
    for c in stall: tot[c]+=int(r[idx[c]] or 0)
    inst.append((n,r[idx['Source']].strip()[:75],{c:int(r[idx[c]] or 0) for c in stall if int(r[idx[c]] or 0)>0}))
T=sum(x[0] for x in inst)
print('total',T); print(', '.join(f'{k[6:]} {100*v/T:.1f}%' for k,v in tot.most_common(12)))
for n,src,st in sorted(inst,key=lambda x:-x[0])[:int(sys.argv[2]) if len(sys.argv)>2 else 20]:
    print(f'{100*n/T:5.1f}% {src:75s}', sorted(st.items(),key=lambda x:-x[1])[:2])
