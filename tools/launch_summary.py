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


def engine_steps(ops):
  """The engine's launch list (csrc/dvb_cnn.cu Plan): average pool moved behind its 1x1 convolution, 1x1 convolutions that
  read the same tensor merged into one GEMM at the position of the first.  Returns [(label, kind, [ops])]."""
  ops = list(ops)
  i = 0
  while i + 1 < len(ops):
    pl, cv = ops[i], ops[i + 1]
    if pl.kind == 'avgpool' and cv.kind == 'conv' and cv.kh == 1 and cv.kw == 1 and cv.src == pl.dst:
      ops[i], ops[i + 1] = cv, pl
      cv._src_override = pl.src if hasattr(cv, '_src_override') is False else cv._src_override
      i += 1
    i += 1
  def src_of(o):
    return getattr(o, '_src_override', o.src)
  groups = {}
  for k, o in enumerate(ops):
    if k > 0 and o.kind == 'conv' and o.kh == 1 and o.kw == 1 and o.stride == 1:
      groups.setdefault(src_of(o), []).append(k)
  merged_into = {}
  for ks in groups.values():
    if 2 <= len(ks) <= 4:
      for k in ks[1:]:
        merged_into[k] = ks[0]
  steps = []
  for k, o in enumerate(ops):
    if k in merged_into:
      continue
    members = [ops[j] for j in groups.get(src_of(o), []) if merged_into.get(j, j) == k] if (o.kind == 'conv' and o.kh == 1 and o.kw == 1 and o.stride == 1 and k > 0) else []
    members = members if len(members) >= 2 else [o]
    steps.append(members)
  return steps, [[o] for o in ops]


steps, unmerged = engine_steps(ops)
n_launch_expected = len(steps) + (1 if fused else 2)
if len(fw) == n_launch_expected - 1:   # conv_rows_kernel: the max pool behind conv3 runs in conv3's epilogue
  for i in range(len(steps) - 1):
    if steps[i][0].kind == 'conv' and steps[i + 1][0].kind == 'maxpool' and steps[i + 1][0].src == steps[i][0].dst and steps[i][0].dst == 's3':
      steps[i:i + 2] = [[steps[i][0], steps[i + 1][0]]]
      break
  n_launch_expected -= 1
if len(fw) != n_launch_expected:   # engine built without the graph rewrites: one launch per op
  steps = unmerged
assert len(fw) == len(steps) + (1 if fused else 2), (len(fw), len(steps))
hw = {'input': (H, W)}
total_us = sum(r[2] for r in fw)
enc = [r for r in rows if r[0] == 'dvb_encode_kernel']
print(f'| # | op | kernel | grid | out HxW | Cin->Cout k | us | GFLOP | TFLOP/s | % of forward |')
print('|---|---|---|---|---|---|---|---|---|---|')
if not fused:
  print(f'| 0 | preprocess+im2col | {fw[0][0]} | {fw[0][1]} | | | {fw[0][2]:.1f} | | | {100 * fw[0][2] / total_us:.1f} |')
by_kernel = {}
flops_total = 0.0
for i, (members, r) in enumerate(zip(steps, fw[0:-1] if fused else fw[1:-1]), 1):
  gf = 0.0
  for o in members:   # shapes follow the ORIGINAL graph: a pool moved behind its conv does not change H x W
    h, w = hw[o.src] if o.src in hw else hw[getattr(o, '_src_override', o.src)]
    oh, ow = modeling.out_hw(o, h, w)
    hw[o.dst] = (oh, ow)
    if o.kind == 'conv':
      gf += 2.0 * batch * oh * ow * o.cout * o.kh * o.kw * o.cin / 1e9
  o = members[0]
  flops_total += gf
  tf = gf / (r[2] * 1e-6) / 1e3 if gf else 0.0
  by_kernel.setdefault(r[0], [0.0, 0.0])
  by_kernel[r[0]][0] += r[2]
  by_kernel[r[0]][1] += gf
  label = ' + '.join(m.name for m in members) if len(members) > 1 else f'{o.name} {o.src}->{o.dst}'
  if len(members) == 2 and members[1].kind == 'maxpool':
    oh, ow = modeling.out_hw(members[0], *hw[members[0].src])
  shape = f'{o.cin}->{"+".join(str(m.cout) for m in members)} {o.kh}x{o.kw}/{o.stride}'
  print(f'| {i} | {label} | {r[0]} | {r[1]} | {oh}x{ow} | {shape} | {r[2]:.1f} | {gf:.1f} | {tf:.0f} | {100 * r[2] / total_us:.1f} |')
print(f'| {len(fw) - 1} | GAP+Dense+softmax | tail_kernel | {fw[-1][1]} | | | {fw[-1][2]:.1f} | | | {100 * fw[-1][2] / total_us:.1f} |')
print()
print(f'forward of {batch} images under ncu (serialised, cold caches): {total_us / 1e3:.2f} ms, {flops_total / (total_us * 1e-6) / 1e3:.0f} TFLOP/s overall')
for k, (us, gf) in sorted(by_kernel.items(), key=lambda kv: -kv[1][0]):
  print(f'- {k}: {us / 1e3:.2f} ms = {100 * us / total_us:.1f} % of the forward' + (f', {gf / (us * 1e-6) / 1e3:.0f} TFLOP/s' if gf else ''))
if enc:
  print(f'- dvb_encode_kernel: {enc[-1][2] / 1e3:.3f} ms per launch = {100 * enc[-1][2] / (total_us + enc[-1][2]):.1f} % of encode + forward')
