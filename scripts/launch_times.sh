# per-kernel device times of one time_stages run (ncu, serialised, cold-cache)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lt_$1.csv python scripts/time_stages.py 32 16 > gpurun_out/lt_$1.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/lt_$1.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',','')); v = v/1e3 if r[ui]=='ns' else (v*1e3 if r[ui]=='ms' else v)
    agg[r[ki].split('(')[0][-40:]].append(v)
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])):
    if k.startswith('void at::') or 'at::' in k: continue
    print('%-42s n=%3d  avg %8.1f us  last %8.1f us'%(k,len(v),sum(v)/len(v),v[-1]))
PY
