"""Per-layer table from an ncu launch list (gpu__time_duration.sum CSV) of one classifier forward.

usage: python tools/launch_summary.py launches.csv BATCH [H W C]  ->  markdown on stdout
The launch order of one forward is: stem_patch_kernel, then one launch per op of modeling.inception_v3_graph
(conv_gemm_kernel / conv_halo_kernel / pool3x3_kernel), then tail_kernel."""
import csv
import sys


def short_name(full):
  """`void <unnamed>::conv_gemm_kernel<0>(CUtensorMap_st, ...)` -> conv_gemm_kernel<0>"""
  head = full.split('(')[0]
  return head.split('::')[-1].replace('void ', '').strip()


sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from deepvariant_b200 import modeling  # noqa: E402

path, batch = sys.argv[1], int(sys.argv[2])
H, W, C = (int(x) for x in sys.argv[3:6]) if len(sys.argv) >= 6 else (100, 221, 7)
rows = []
with open(path) as f:
  for r in csv.reader(f):
    if len(r) >= 15 and r[0].isdigit() and r[12] == 'gpu__time_duration.sum':
      name = r[4]
      short = short_name(name).split('<')[0]
      rows.append((short, r[8], float(r[14]) / 1e3))   # us
# find the LAST complete forward: ... stem_patch ... tail
tails = [i for i, r in enumerate(rows) if r[0] == 'tail_kernel']
stems = [i for i, r in enumerate(rows) if r[0].startswith('stem_')]
end = tails[-1]
start = max(i for i in stems if i < end)
fw = rows[start:end + 1]
ops, _ = modeling.inception_v3_graph(C)
fused = fw[0][0] == 'stem_conv1_kernel'   # preprocess + im2col + conv1 in one launch
assert len(fw) == len(ops) + (1 if fused else 2), (len(fw), len(ops))
hw = {'input': (H, W)}
total_us = sum(r[2] for r in fw)
enc = [r for r in rows if r[0] == 'dvb_encode_kernel']
print(f'| # | op | kernel | grid | out HxW | Cin->Cout k | us | GFLOP | TFLOP/s | % of forward |')
print('|---|---|---|---|---|---|---|---|---|---|')
if not fused:
  print(f'| 0 | preprocess+im2col | {fw[0][0]} | {fw[0][1]} | | | {fw[0][2]:.1f} | | | {100 * fw[0][2] / total_us:.1f} |')
by_kernel = {}
flops_total = 0.0
for i, (o, r) in enumerate(zip(ops, fw[0:-1] if fused else fw[1:-1]), 1):
  h, w = hw[o.src]
  oh, ow = modeling.out_hw(o, h, w)
  hw[o.dst] = (oh, ow)
  gf = 2.0 * batch * oh * ow * o.cout * o.kh * o.kw * o.cin / 1e9 if o.kind == 'conv' else 0.0
  flops_total += gf
  tf = gf / r[2] / 1e3 * 1e3 if gf else 0.0   # GFLOP / us = PFLOP/s*1e-3 -> TFLOP/s = gf / us * 1e3 / 1e3
  tf = gf / (r[2] * 1e-6) / 1e3 if gf else 0.0
  by_kernel.setdefault(r[0], [0.0, 0.0])
  by_kernel[r[0]][0] += r[2]
  by_kernel[r[0]][1] += gf
  print(f'| {i} | {o.name} {o.src}->{o.dst} | {r[0]} | {r[1]} | {oh}x{ow} | {o.cin}->{o.cout} {o.kh}x{o.kw}/{o.stride} | {r[2]:.1f} | '
        f'{gf:.1f} | {tf:.0f} | {100 * r[2] / total_us:.1f} |')
print(f'| {len(fw) - 1} | GAP+Dense+softmax | tail_kernel | {fw[-1][1]} | | | {fw[-1][2]:.1f} | | | {100 * fw[-1][2] / total_us:.1f} |')
print()
print(f'forward of {batch} images under ncu (serialised, cold caches): {total_us / 1e3:.2f} ms, {flops_total / (total_us * 1e-6) / 1e3:.0f} TFLOP/s overall')
for k, (us, gf) in sorted(by_kernel.items(), key=lambda kv: -kv[1][0]):
  print(f'- {k}: {us / 1e3:.2f} ms = {100 * us / total_us:.1f} % of the forward' + (f', {gf / (us * 1e-6) / 1e3:.0f} TFLOP/s' if gf else ''))
if enc:
  print(f'- dvb_encode_kernel: {enc[-1][2] / 1e3:.3f} ms per launch = {100 * enc[-1][2] / (total_us + enc[-1][2]):.1f} % of encode + forward')
