python tools/prof_boot.py 1e8 512
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/boot_launches.csv python tools/prof_boot.py 1e8 64 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/boot_launches.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[iv].replace(",",""))
    except: continue
    k=r[ik].split("(")[0][:60]
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print("%-62s x%-5d %10.2f ms"%(k,c,v/1e6))
PY
