#!/bin/bash
# Round 2, GPU call 46: L2 prefetch of the rows' CIGAR / bases / qualities in phase A (variant pf) against the shipped build; suites on it.
mkdir -p gpurun_out
for rep in 1 2; do for lay in wgs pacbio; do for v in shipped dyn; do
  if [ $v = shipped ]; then unset DVB_LIB_PATH; else export DVB_LIB_PATH=$PWD/_variants/libdvb_$v.so; fi
  extra=""; [ $lay = pacbio ] && extra="--pacbio"
  timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 $extra > gpurun_out/c46_enc_${lay}_${v}_$rep.json 2>/dev/null; echo "$lay $v $rep: $(cut -c1-130 gpurun_out/c46_enc_${lay}_${v}_$rep.json)"
done; done; done
export DVB_LIB_PATH=$PWD/_variants/libdvb_dyn.so
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py tests/test_channel_planes.py tests/test_pair_support.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c46_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/c46_pytest.log
