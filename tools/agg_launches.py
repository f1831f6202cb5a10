"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for d in csv.DictReader(lines):
    try:
        v = float(d['Metric Value'].replace(',', ''))
    except (KeyError, ValueError):
        continue
    us = v / 1e3 if d['Metric Unit'] in ('ns', 'nsecond') else v
    n = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('dx::', '').replace('<unnamed>::', '')
    agg[n][0] += 1
    agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print(f'total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches')
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f'{t:9.1f} us {100 * t / tot:5.1f}% {c:4d}  {n[:100]}')
