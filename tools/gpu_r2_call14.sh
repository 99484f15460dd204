#!/bin/bash
# Round 2, GPU call 14 (2 GPUs): the bench line under torchrun at N = 2 (new blocks take part in barriers), the reference arm's ranks,
# and the fused multi-GPU CLI on the planted genome with 4 tasks on 2 devices.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c14_bench_n2.json 2> gpurun_out/c14_bench_n2.err; echo "bench n2 exit $?"; cat gpurun_out/c14_bench_n2.json | cut -c1-900; tail -3 gpurun_out/c14_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c14_bench_ref_n2.json 2> gpurun_out/c14_bench_ref_n2.err; echo "bench ref n2 exit $?"; cut -c1-400 gpurun_out/c14_bench_ref_n2.json
timeout 600 python - <<'PY' > gpurun_out/c14_cli_2gpu.log 2>&1
import os, sys, time, tempfile, pathlib
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_candidates as tc
from deepvariant_b200 import cli, tfrecord
tmp = pathlib.Path(tempfile.mkdtemp())
fa, bam_path, genome, sites = tc._planted_case(tmp)
outs = {}
for name, extra in (('one', ['--num_shards', '1']), ('four_on_two', ['--num_shards', '4', '--num_gpus', '2'])):
  d = str(tmp / name)
  t0 = time.time()
  assert cli.run_deepvariant(['--model_type', 'WGS', '--ref', fa, '--reads', bam_path, '--regions', 'chr20:1001-5000', '--customized_model', 'random:3',
                              '--output_dir', d, '--output_vcf', os.path.join(d, 'o.vcf')] + extra) == 0
  outs[name] = open(os.path.join(d, 'o.vcf')).read()
  print(name, 'seconds', round(time.time() - t0, 1), 'cvo shards', len(tfrecord.resolve_input_paths(os.path.join(d, 'call_variants_output.tfrecord.gz'))))
assert outs['one'] == outs['four_on_two'], 'VCF differs between 1 task and 4 tasks on 2 GPUs'
print('VCF identical:', len(outs['one'].splitlines()), 'lines')
PY
echo "cli exit $?"; tail -5 gpurun_out/c14_cli_2gpu.log
