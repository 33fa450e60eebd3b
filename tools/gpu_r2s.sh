mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fc_rows|fc_cols|se_fc_bwd|hsig|dgate|wgrad_reduce|wgrad_tc" --csv --log-file gpurun_out/small_ncu.csv python tools/small_bench.py > gpurun_out/small_ncu.log 2>&1; tail -2 gpurun_out/small_ncu.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/small_ncu.csv', errors='ignore')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[start+1:]:
    if len(r)<=vi: continue
    key=(r[ki].split('(')[0][-40:], r[gi])
    agg.setdefault(key, []).append(float(r[vi].replace(',',''))/1000)
for k,v in agg.items():
    v=sorted(v); print(f"{k[0]:42s} {k[1]:16s} n={len(v):4d} median {v[len(v)//2]:7.1f} us  min {v[0]:7.1f}")
PY
