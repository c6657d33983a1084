import csv,sys,collections,re
rows=list(csv.reader(open(sys.argv[1],errors='ignore')))
hdr=None;agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r;continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    n=re.sub(r'\(.*','',d['Kernel Name'])
    v=float(d['Metric Value'].replace(',',''))
    u=d['Metric Unit']
    if u in('ns','nsecond'): v/=1e3
    elif u in ('ms','msecond'): v*=1e3
    agg[n][0]+=1;agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
for n,(c,t) in sorted(agg.items(),key=lambda x:-x[1][1])[:30]:
    print(f'{t/1e3:9.2f} ms {100*t/tot:5.1f}% n={c:5d} avg={t/c:8.1f}us {n[:90]}')
print('total',tot/1e3)
