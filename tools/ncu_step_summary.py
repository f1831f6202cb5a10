"""Per-kernel summary of an ncu report holding EVERY launch of one training step (tools/profile_step_all.sh):
   python tools/ncu_step_summary.py gpurun_out/r2_step_all.ncu-rep profiles/r2_step_kernels.json [profiles/r2_step_kernels.md]
For each kernel name: launches, total / mean duration, share of the step, DRAM bytes (read + write) and achieved HBM GB/s,
tensor-pipe active %, issue-slot active %, and the two largest warp-stall reasons (cycles per issued instruction)."""
import collections, csv, json, re, subprocess, sys

rep, out_json = sys.argv[1], sys.argv[2]
out_md = sys.argv[3] if len(sys.argv) > 3 else None
HBM_PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if __import__('os').path.exists('MEASURED_PEAKS.json') else 6650.0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}


def val(r, name, scale=True):
    i = col.get(name)
    if i is None or r[i] in ('', 'n/a'):
        return None
    v = float(r[i].replace(',', ''))
    return v * MULT.get(units[i], 1) if scale else v


stall_cols = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')]
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('dx::', '').replace('<unnamed>::', '').replace('unnamed>::', '')
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, tensor=0.0, issue=0.0, stalls=collections.Counter(), regs=None, grid=None, block=None))
    us = val(r, 'gpu__time_duration.sum') or 0.0
    a['n'] += 1
    a['us'] += us
    a['rd'] += val(r, 'dram__bytes_read.sum') or 0.0
    a['wr'] += val(r, 'dram__bytes_write.sum') or 0.0
    a['tensor'] += us * (val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', False) or 0.0)
    a['issue'] += us * (val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active', False) or 0.0)
    for h in stall_cols:
        v = val(r, h, False)
        if v:
            a['stalls'][h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', '').replace('.ratio', '')] += us * v
    a['regs'] = val(r, 'launch__registers_per_thread', False)
    a['grid'], a['block'] = val(r, 'launch__grid_size', False), val(r, 'launch__block_size', False)
total = sum(a['us'] for a in agg.values())
kernels = []
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    us = a['us']
    top = [(k, v / us) for k, v in a['stalls'].most_common(2)] if us else []
    kernels.append({'kernel': name, 'launches': a['n'], 'total_us': round(us, 1), 'mean_us': round(us / a['n'], 2), 'share_pct': round(100 * us / total, 2),
                    'dram_read_mb': round(a['rd'] / 1e6, 1), 'dram_write_mb': round(a['wr'] / 1e6, 1),
                    'hbm_gbs': round((a['rd'] + a['wr']) / us / 1e3, 1) if us else None,
                    'hbm_frac_of_measured': round((a['rd'] + a['wr']) / us / 1e3 / HBM_PEAK, 3) if us else None,
                    'tensor_pipe_active_pct': round(a['tensor'] / us, 1) if us else None, 'issue_active_pct': round(a['issue'] / us, 1) if us else None,
                    'top_stalls': [(k, round(v, 2)) for k, v in top], 'registers': a['regs'], 'grid': a['grid'], 'block': a['block']})
summary = {'source': f'ncu --section SpeedOfLight/MemoryWorkloadAnalysis/WarpStateStats/... --clock-control none, every launch of one eager training step ({rep}); '
                     'per-launch times are cold-cache and serialised: compare SHARES', 'hbm_peak_gbs_measured': HBM_PEAK,
           'total_us': round(total, 1), 'launches': sum(a['n'] for a in agg.values()), 'kernels': kernels}
json.dump(summary, open(out_json, 'w'), indent=1)
lines = [f'total {total:.1f} us over {summary["launches"]} launches (serialised, cold cache)', '',
         '| kernel | launches | total us | share | mean us | DRAM rd+wr MB | HBM GB/s (frac of measured) | tensor pipe % | issue active % | top stalls (cycles/issue) |',
         '|---|---|---|---|---|---|---|---|---|---|']
for k in kernels[:45]:
    lines.append(f"| `{k['kernel'][:70]}` | {k['launches']} | {k['total_us']} | {k['share_pct']} % | {k['mean_us']} | {k['dram_read_mb'] + k['dram_write_mb']:.0f} | "
                 f"{k['hbm_gbs']} ({k['hbm_frac_of_measured']}) | {k['tensor_pipe_active_pct']} | {k['issue_active_pct']} | {k['top_stalls']} |")
if out_md:
    open(out_md, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:40]))
