for mode in single two; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/boot_$mode.csv python tools/prof_boot.py 1e8 128 $mode > /dev/null 2>&1
python - $mode <<'PY'
import csv, collections, sys
rows=[r for r in csv.reader(open(f"gpurun_out/boot_{sys.argv[1]}.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    try: v=float(r[iv].replace(",",""))
    except: continue
    k=r[ik].split("(")[0][:60]
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v; tot+=v
print(sys.argv[1], "total %.1f ms (2 runs of B=128)"%(tot/1e6))
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:7]: print("   %-58s x%-5d %10.2f ms"%(k,c,v/1e6))
PY
done
