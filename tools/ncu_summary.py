"""Compact markdown table from `ncu -i X.ncu-rep --page raw --csv` output (one row per launch).
usage: python tools/ncu_summary.py raw.csv > profiles/summary.md"""
import csv
import sys


def short_name(full):
  """`void <unnamed>::conv_gemm_kernel<0>(CUtensorMap_st, ...)` -> conv_gemm_kernel<0>"""
  head = full.split('(')[0]
  return head.split('::')[-1].replace('void ', '').strip()


rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
COLS = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'), ('launch__registers_per_thread', 'regs')]
cols = [(k, n) for k, n in COLS if k in ix]
print('| # | kernel | grid | ' + ' | '.join(f'{n} ({units[ix[k]]})' if units[ix[k]] not in ('%', '') else n for k, n in cols) + ' |')
print('|---|---|---|' + '---|' * len(cols))
tot_t = 0.0
tc_w = 0.0
by = {}
for i, r in enumerate(data):
  name = short_name(r[ix['Kernel Name']])
  vals = []
  for k, _ in cols:
    try:
      vals.append(f'{float(r[ix[k]]):.3g}')
    except ValueError:
      vals.append(r[ix[k]])
  print(f'| {i} | {name} | {r[ix["Grid Size"]]} | ' + ' | '.join(vals) + ' |')
  t = float(r[ix['gpu__time_duration.sum']])
  tot_t += t
  k = 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active'
  if k in ix:
    tc_w += t * float(r[ix[k]])
  b = by.setdefault(name, [0.0, 0.0, 0.0, 0])
  b[0] += t
  b[1] += float(r[ix['dram__bytes_read.sum']])
  b[2] += float(r[ix['dram__bytes_write.sum']])
  b[3] += 1
print()
tu = units[ix['gpu__time_duration.sum']]
print(f'total {tot_t:.3f} {tu} over {len(data)} launches; time-weighted tensor-pipe utilisation {tc_w / tot_t:.1f} %')
for name, (t, rd, wr, n) in sorted(by.items(), key=lambda kv: -kv[1][0]):
  print(f'- {name}: {n} launches, {t:.3f} {tu} ({100 * t / tot_t:.1f} %), DRAM read {rd:.3g} + write {wr:.3g} (same unit as the columns)')
