import csv, collections, re, sys
fn = sys.argv[1]
with open(fn) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def name(n):
    n=re.sub(r'\(.*','',n); return n[:80]
def us(row):
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    return v/1e3 if u=='ns' else v*1e3 if u=='ms' else v*1e6 if u=='s' else v
agg=collections.OrderedDict()
for row in rows:
    k=name(row['Kernel Name']); agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=us(row)
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 18]:
    print(f"{t:10.1f} us {n:4d}  {100*t/tot:5.1f}%  {k}")
print('total us', round(tot,1), 'launches', len(rows))
if len(sys.argv)>3:
    for row in rows:
        if re.search(sys.argv[3], row['Kernel Name']):
            print(row['ID'], name(row['Kernel Name'])[:36], row['Grid Size'], round(us(row),1))
