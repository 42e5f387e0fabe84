import csv,sys
from collections import defaultdict
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    name=r[ik].split('(')[0][-60:]
    try: v=float(r[iv].replace(',',''))
    except: continue
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v for _,v in agg.values())
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 14]: print(f"{k:62s} {c:5d} {v/1e3:11.1f} us {100*v/tot:5.1f}%  avg {v/c/1e3:9.1f}")
