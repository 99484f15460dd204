"""Copies the reference's small CRAM 3.0 test vectors into tests/golden/cram/ (third_party/nucleus/testdata: the same three reads written
with and without an embedded reference, the SAM they were made from, and the FASTA they were written against; sam_test.py:250-262,
sam_writer_test.cc:300-330 use them).  Run in the build container (needs /root/reference)."""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/third_party/nucleus/testdata'
for name in ('test_cram.embed_ref_0_version_3.0.cram', 'test_cram.embed_ref_1_version_3.0.cram', 'test_cram.sam', 'test.fasta', 'test.fasta.fai'):
  shutil.copyfile(os.path.join(SRC, name), os.path.join(ROOT, 'tests/golden/cram', name))
  os.chmod(os.path.join(ROOT, 'tests/golden/cram', name), 0o644)
