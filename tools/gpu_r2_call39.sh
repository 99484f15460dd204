#!/bin/bash
# Round 2, GPU call 39: the new encoder defaults (MIN_BLOCKS 2 + word loads; the per-base-extras instantiation stays at 4) - WGS and
# PACBIO layouts against the old defaults (prebuilt), the whole GPU suite, the bench line.
mkdir -p gpurun_out
for rep in 1 2; do for lay in wgs pacbio; do for v in new old; do
  if [ $v = new ]; then unset DVB_LIB_PATH; else export DVB_LIB_PATH=$PWD/_variants/libdvb_t256.so; fi
  extra=""; [ $lay = pacbio ] && extra="--pacbio"
  timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 $extra > gpurun_out/c39_enc_${lay}_${v}_$rep.json 2>/dev/null; echo "$lay $v $rep: $(cut -c1-130 gpurun_out/c39_enc_${lay}_${v}_$rep.json)"
done; done; done
unset DVB_LIB_PATH
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c39_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/c39_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c39_bench_n1.json 2> gpurun_out/c39_bench_err.txt; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/c39_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_encoder'], d['config5_encode_only'], d['parity'])"
