"""profiles/r2_stress_dram.json from an ncu launch list (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) of ONE
eager training step of the stress config:  python tools/stress_dram.py gpurun_out/r2_stress_launches.csv profiles/r2_stress_dram.json"""
import collections, csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}
per = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot_b = tot_us = 0.0
ids = set()
for d in csv.DictReader(lines):
    try:
        v = float(d['Metric Value'].replace(',', '')) * MULT.get(d['Metric Unit'], 1)
    except (KeyError, ValueError):
        continue
    n = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('dx::', '').replace('<unnamed>::', '')
    if d['Metric Name'].startswith('dram__bytes'):
        per[n][1] += v
        tot_b += v
    elif d['Metric Name'].startswith('gpu__time_duration'):
        per[n][2] += v
        tot_us += v
        per[n][0] += 1
        ids.add(d['ID'])
top = sorted(per.items(), key=lambda kv: -kv[1][1])[:12]
out = {'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over ONE eager training step of '
                 'bench.py --config stress (B=128, T<=1500), tools/step_launches.py', 'launches': len(ids), 'dram_bytes_per_step': tot_b,
       'serialised_us_per_step': tot_us, 'top_kernels_by_dram_bytes': [{'kernel': k[:80], 'launches': v[0], 'dram_gb': round(v[1] / 1e9, 3), 'us': round(v[2], 1)} for k, v in top]}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != 'top_kernels_by_dram_bytes'}))
