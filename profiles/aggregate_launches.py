import csv,re,collections,sys
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict(); tot=0; n=0
for row in r:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit in ('nsecond','ns'): v/=1e3
    elif unit=='msecond': v*=1e3
    name=re.sub(r'swg::','',name)
    m=re.search(r'k_for<(\w+)',name)
    if m: name='k_for<'+m.group(1)+'>'
    else:
        m=re.search(r'(sc_\w+)<.*?(\w+)\(.*lambda.*?#(\d+)',name)
        if m: name=f"{m.group(1)}<{m.group(2)} #{m.group(3)}>"
        else: name=re.sub(r'\(.*','',name)
    agg.setdefault(name[:70],[0,0]); agg[name[:70]][0]+=v; agg[name[:70]][1]+=1; tot+=v; n+=1
print("total us",round(tot,1), "launches", n)
for k,(v,c) in sorted(agg.items(), key=lambda x:-x[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print(f"{v:10.1f} us {c:4d}  {100*v/tot:5.1f}%  {k}")
