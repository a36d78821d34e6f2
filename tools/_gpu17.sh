timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -p no:cacheprovider -k "bracket or percentile or select or stats" 2>&1 | tail -3
python tools/prof_select.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sel_launches.csv python tools/prof_select.py 1e8 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/sel_launches.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
names=[(r[ik][:50], float(r[iv].replace(",",""))) for r in rows[1:]]
idx=[i for i,(k,_) in enumerate(names) if "bracket_pass" in k]
if idx:
    i=idx[-1]; j=i
    while j>0 and "select_sample" not in names[j][0]: j-=1
    end=i
    while end+1<len(names) and "select_finish" not in names[end][0]: end+=1
    from collections import OrderedDict
    agg=OrderedDict(); tot=0
    for k,v in names[j:end+1]:
        agg.setdefault(k,[0,0]); agg[k][0]+=1; agg[k][1]+=v; tot+=v
    for k,(c,v) in agg.items(): print("%-52s x%-3d %9.1f us"%(k,c,v/1e3))
    print("sum of kernel times (n=1e8, cold-cache, serialized): %.1f us over %d launches"%(tot/1e3,end-j+1))
PY
