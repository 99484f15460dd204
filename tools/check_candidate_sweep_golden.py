"""Pins `make_examples --mode candidate_sweep` against the reference's golden.candidate_positions (+ its 3 shards;
scripts/create_golden.sh:202-213): the int32 stream of candidate positions, END_OF_PARTITION after every 1-kb partition and
END_OF_REGION at the end of the calling region, written by the stage CLI from BAM + FASTA, byte for byte.
Writes tests/golden/candidate_sweep_report.json and copies the four small goldens into tests/golden/ for the merge tests."""
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepvariant_b200 import cli  # noqa: E402

T = '/root/reference/deepvariant/testdata/'


def main():
  args = ['--mode', 'candidate_sweep', '--ref', T + 'input/ucsc.hg19.chr20.unittest.fasta.gz', '--reads', T + 'input/NA12878_S1.chr20.10_10p1mb.bam',
          '--regions', 'chr20:10,000,000-10,010,000', '--channel_list', 'BASE_CHANNELS,insert_size']
  report = {}
  with tempfile.TemporaryDirectory() as tmp:
    cli.make_examples(args + ['--examples', os.path.join(tmp, 'e.tfrecord.gz'), '--candidate_positions', os.path.join(tmp, 'positions')])
    report['unsharded_byte_identical'] = open(os.path.join(tmp, 'positions'), 'rb').read() == open(T + 'golden.candidate_positions', 'rb').read()
    shards = []
    for task in range(3):
      cli.make_examples(args + ['--examples', os.path.join(tmp, 'e.tfrecord@3.gz'), '--candidate_positions', os.path.join(tmp, 'p@3'), '--task', str(task)])
      shards.append(open(os.path.join(tmp, f'p-0000{task}-of-00003'), 'rb').read() == open(T + f'golden.candidate_positions-0000{task}-of-00003', 'rb').read())
    report['shards_byte_identical'] = shards
  for name in ['golden.candidate_positions'] + [f'golden.candidate_positions-0000{i}-of-00003' for i in range(3)]:
    shutil.copy(T + name, os.path.join(ROOT, 'tests/golden', name))
    os.chmod(os.path.join(ROOT, 'tests/golden', name), 0o644)
  with open(os.path.join(ROOT, 'tests/golden/candidate_sweep_report.json'), 'w') as f:
    json.dump(report, f, indent=1)
  print(json.dumps(report))


if __name__ == '__main__':
  main()
