# launch list of the 20 M pile call (which kernels the inversion stage spends its time in)
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches57_pile20m.csv python profiles/bench_skew.py 20000000 100000 2>&1 | tail -2
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches57_pile20m.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[iv].replace(',','')); u=r[iu]
    v = v/1e3 if u in ('ns','nsecond') else v if u in ('us','usecond') else v*1e3 if u in ('ms','msecond') else v
    k=r[ik][:90]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('total us',tot)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{a[1]:12.1f} us {a[0]:5d}  {k}")
PY
