timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
python bench.py --configs stats --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
s=j['configs']['stats']
print('percentiles', s['percentiles'], 'moments', s['moments']['s'], 'boot', s['value'])
print('headline', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['peak'], j['roofline']['peak_source'])
"
