"""Hot SASS instructions of a kernel from `ncu -i X.ncu-rep --page source --csv`: samples, dominant stall reasons.
usage: python tools/ncu_hot.py rep.ncu-rep [kernel_index] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
kernels, cur = [], None
for row in csv.reader(io.StringIO(txt)):
  if row and row[0] == 'Kernel Name':
    cur = {'name': row[1], 'hdr': None, 'rows': []}
    kernels.append(cur)
  elif cur is not None and row and row[0] == 'Address':
    cur['hdr'] = row
  elif cur is not None and cur['hdr'] and len(row) >= len(cur['hdr']) - 2:
    cur['rows'].append(row)
k = kernels[which]
h = k['hdr']
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[ix['# Samples']] or 0) for r in k['rows'])
print(k['name'][:90], 'launch', which, 'total samples', tot, 'instructions', len(k['rows']))
agg = {s: sum(int(r[ix[s]] or 0) for r in k['rows']) for s in stalls}
print('stall totals:', ', '.join(f'{s[6:]} {100*v/max(1,sum(agg.values())):.1f}%' for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
rows = sorted(enumerate(k['rows']), key=lambda ir: -int(ir[1][ix['# Samples']] or 0))[:top]
for i, r in sorted(rows):
  n = int(r[ix['# Samples']] or 0)
  dom = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
  print(f'{i:5d} {100*n/max(1,tot):5.1f}% {r[ix["Source"]].strip()[:70]:70s} exec {r[ix["Instructions Executed"]]:>8s}  {dom[0][1]} {dom[0][0]}, {dom[1][1]} {dom[1][0]}')
