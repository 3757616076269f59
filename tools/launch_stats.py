#!/usr/bin/env python
"""Per-kernel time / instruction series from an ncu launch list (dev helper).  usage: launch_stats.py file.csv kernel_substr [stride]"""
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]
ki,mi,vi,idi=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("ID")
recs={}
for r in rows[1:]:
    recs.setdefault(r[idi],{"k":r[ki].split("(")[0]})[r[mi]]=float(r[vi].replace(",",""))
seq=[v for k,v in sorted(recs.items(), key=lambda x:int(x[0]))]
stride=int(sys.argv[3]) if len(sys.argv)>3 else 1
games=float(sys.argv[4]) if len(sys.argv)>4 else 16384.0
for sub in sys.argv[2].split(","):
    a=[v for v in seq if sub in v['k']][0::stride]
    if not a: continue
    print(sub,"time us:",[round(x['gpu__time_duration.sum']/1000) for x in a][:90])
    if 'smsp__inst_executed.sum' in a[0]:
        print(sub,"instr/unit:",[round(x['smsp__inst_executed.sum']/games) for x in a][:90])
        print(sub,"mean us %.1f mean instr %.0f"%(sum(x['gpu__time_duration.sum'] for x in a)/len(a)/1000, sum(x['smsp__inst_executed.sum'] for x in a)/len(a)/games))
