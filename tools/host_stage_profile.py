"""Host-side time of the fused make_examples flow WITHOUT a GPU (development aid): the encoder / classifier calls are replaced by
stubs that drop the packed batches, everything before them (BAM decode, realigner, allele counter + caller, region packer, allele
keys) runs as in the product.  Prints the --runtime_by_region sums and the top of a cProfile.
usage: python tools/host_stage_profile.py [--regions chr20:10,000,001-10,010,000] [--norealign]"""
import argparse
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepvariant_b200 import call_variants as cv, cli, fused, make_examples_native as men, pileup_image as pi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--regions', default='chr20:10,000,001-10,010,000')
ap.add_argument('--norealign', action='store_true')
ap.add_argument('--top', type=int, default=35)
a = ap.parse_args()


class StubCnn:
  def __init__(self, *args, **kw):
    pass
  def close(self):
    pass


class StubEncoder:
  def __init__(self, params):
    self.params = params
    self.shape = (params.height, params.width, params.num_channels + params.num_alt_channels)


cv.GpuCnn = StubCnn
cv.load_weights = lambda *args, **kw: None
cv.model_example_info = lambda *args, **kw: None
cv.check_example_info = lambda *args, **kw: None
men.ExamplesGenerator._gpu = lambda self: StubEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height))
n_images = [0]


def flush(self):
  n_images[0] += self._pending
  self._packed, self._images, self._order, self._pending = [], [], [], 0


fused.FusedCaller.flush = flush
fused.FusedCaller.close = lambda self: (self.flush(), 0)[1]
g = os.path.join(ROOT, 'tests', 'golden')
with tempfile.TemporaryDirectory() as d:
  args = ['--mode', 'calling', '--ref', os.path.join(g, 'quickstart.chr20_10mb.fa.gz'), '--reads', os.path.join(g, 'quickstart.chr20_10mb.bam'), '--regions', a.regions,
          '--call_variants_outfile', os.path.join(d, 'cvo@1.tfrecord.gz'), '--checkpoint', 'random:1', '--channel_list', 'BASE_CHANNELS,insert_size',
          '--runtime_by_region', os.path.join(d, 'rt.tsv')] + (['--norealign_reads'] if a.norealign else [])
  pr = cProfile.Profile()
  t0 = time.time()
  pr.enable()
  cli.make_examples(args)
  pr.disable()
  dt = time.time() - t0
  rows = [line.rstrip('\n').split('\t') for line in open(os.path.join(d, 'rt.tsv'))]
  sums = {k: round(sum(float(r[i + 1]) for r in rows[1:]), 2) for i, k in enumerate(rows[0][1:8])}
print(f'wall {dt:.2f} s, images {n_images[0]}, per-stage sums {sums}')
pstats.Stats(pr).sort_stats('cumulative').print_stats(a.top)
