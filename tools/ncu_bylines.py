import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
tiles=int(sys.argv[2]) if len(sys.argv)>2 else 50040
hi=[i for i,r in enumerate(rows) if r and r[0]=='Line No'][0]
hdr=rows[hi]
# columns: Line No, Source, Address, Source(sass), ...
iS=hdr.index('# Samples'); iN=hdr.index('Instructions Executed')
cur=None; agg={}
order=[]
for r in rows[hi+1:]:
    if len(r)<=iN: continue
    if r[0]!='':
        cur=(r[0], r[1][:105]); 
        if cur not in agg: agg[cur]=[0,0]; order.append(cur)
        continue
    try: s=int(r[iS]); n=int(r[iN])
    except: continue
    agg[cur][0]+=s; agg[cur][1]+=n
tot=sum(v[0] for v in agg.values())
for k in order:
    s,n=agg[k]
    if s*100/tot>0.7 or n/tiles>60: print(f'{k[0]:>5} {100*s/tot:5.1f}% {n/tiles:7.1f}  {k[1]}')
