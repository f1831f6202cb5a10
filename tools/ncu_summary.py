"""Summarise an .ncu-rep (ncu --set full) into a small JSON: python tools/ncu_summary.py gpurun_out/prof.ncu-rep out.json [note]"""
import csv, json, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = {'kernel': r[hdr.index('Kernel Name')]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {'value': float(r[i].replace(',', '')), 'unit': units[i]}
    launches.append(d)
def to_bytes(m):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return m['value'] * mult.get(m['unit'], 1)
first = launches[0]
summary = {'source': f'ncu --set full --clock-control none --import-source on ({sys.argv[1]})', 'note': sys.argv[3] if len(sys.argv) > 3 else '',
           'traffic_bytes_per_launch': to_bytes(first['dram__bytes_read.sum']) + to_bytes(first['dram__bytes_write.sum']), 'launches': launches}
json.dump(summary, open(sys.argv[2], 'w'), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != 'launches'}), len(launches), 'launch(es)')
