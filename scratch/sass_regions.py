import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
H=rows[hi]; ii=H.index("Instructions Executed"); si=H.index("Source"); sa=H.index("# Samples")
data=[]
for r in rows[hi+1:]:
    if len(r)>ii and r[0].startswith("0x"):
        data.append((r[si].strip(), int(r[ii] or 0), int(r[sa] or 0)))
tot=sum(d[1] for d in data); ts=sum(d[2] for d in data); print("total inst", tot, "samples", ts, "n sass", len(data))
marker=sys.argv[2] if len(sys.argv)>2 else "BAR"
seg=0; acc=0; sacc=0; start=0
for i,(s,n,sm) in enumerate(data):
    acc+=n; sacc+=sm
    if s.startswith(marker) or i==len(data)-1:
        print("seg %2d sass[%4d..%4d] inst %10d (%5.1f%%) samples %6d (%5.1f%%)  ends with %s"%(seg,start,i,acc,100*acc/tot,sacc,100*sacc/max(ts,1),s[:50]))
        seg+=1; acc=0; sacc=0; start=i+1
