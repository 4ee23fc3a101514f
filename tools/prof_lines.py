import csv, sys
path=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
cur=None; agg={}
hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); continue
    if r[0]=="Function Name" or hdr is None: continue
    if r[0]!="":  # source line summary row
        try: v=float(r[ie]); s=float(r[isamp])
        except: continue
        k=(cur,int(r[0]),r[1][:100])
        a=agg.setdefault(k,[0,0]); a[0]+=v; a[1]+=s
tot=sum(a[0] for a in agg.values()); tots=sum(a[1] for a in agg.values())
print('total warp-inst',tot,'samples',tots)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/tots:5.1f}% samp {k[0]}:{k[1]}: {k[2]}")
