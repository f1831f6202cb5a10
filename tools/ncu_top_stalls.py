"""stdin: `ncu -i X.ncu-rep --page source --csv`  ->  top SASS instructions by warp-stall samples + opcode mix (stdout)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
h = rows[hi]
i_src, i_s, i_e = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
data = [(int(r[i_s] or 0), int(r[i_e] or 0), k, r[i_src].strip()) for k, r in enumerate(rows[hi + 1:]) if len(r) > i_e]
tot = sum(d[0] for d in data) or 1
print(rows[0][1][:160] if len(rows[0]) > 1 else '')
print(f'warp-stall samples {tot}, SASS instructions {len(data)}, warp-level instructions executed {sum(d[1] for d in data)}')
for s, e, k, src in sorted(data, reverse=True)[:25]:
    print(f'{s:7d} {100 * s / tot:5.1f}%  exec {e:9d}  #{k:5d}  {src[:100]}')
ops = {}
for s, e, k, src in data:
    t = src.split()
    op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')).split('.')[0]
    ops[op] = ops.get(op, 0) + e
print('opcode mix (warp-level executed):', sorted(ops.items(), key=lambda kv: -kv[1])[:18])
