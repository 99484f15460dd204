"""Repeats one forward many times per plan and compares every run with the first (bitwise) and with the default plan: a data race in a
kernel shows up as run-to-run differences.  Development aid (GPU)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvariant_b200 import call_variants as cv, modeling  # noqa: E402

PLANS = {
    'default': {},
    'persist2': {'DVB_CNN_PERSIST': '2'},
    'persist2_pair1': {'DVB_CNN_PERSIST': '2', 'DVB_CNN_PAIR': '1'},
    'persist2_pair3': {'DVB_CNN_PERSIST': '2', 'DVB_CNN_PAIR': '3', 'DVB_PERSIST_MIN_N': '128'},
}
KEYS = ('mixed4', 'mixed5', 'mixed6', 'mixed7', 'mixed10')


def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
  batch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
  shape = (100, 221, 7)
  w = modeling.random_weights(7, 33)
  g = torch.Generator().manual_seed(33)
  imgs = torch.randint(0, 255, (batch,) + shape, dtype=torch.uint8, generator=g).numpy()
  out, ref = {}, None
  for name, env in PLANS.items():
    for k in ('DVB_CNN_PERSIST', 'DVB_CNN_PAIR', 'DVB_PERSIST_MIN_N'):
      os.environ.pop(k, None)
    os.environ.update(env)
    net = cv.GpuCnn(w, shape, device=0, max_batch=batch)
    first, n_diff, worst = None, 0, 0.0
    for r in range(reps):
      probs = net.forward_host(imgs)
      t = {k: net.debug_tensor(k, batch) for k in KEYS}
      if first is None:
        first = (probs, t)
        continue
      d = max([float(np.abs(probs - first[0]).max())] + [float(np.abs(t[k] - first[1][k]).max()) for k in KEYS])
      if d > 0:
        n_diff += 1
        worst = max(worst, d)
    if ref is None:
      ref = first
    vs_default = max(float(np.abs(first[1][k] - ref[1][k]).max()) / max(float(np.abs(ref[1][k]).max()), 1e-9) for k in KEYS)
    out[name] = {'reps': reps, 'runs_differing_from_first': n_diff, 'worst_abs_difference': worst, 'first_run_vs_default_relative': vs_default}
    net.close()
    print(name, out[name], flush=True)
  print(json.dumps(out))


if __name__ == '__main__':
  main()
