#!/bin/bash
# Round 2, GPU call 31: encoder A/B - DVB_ENC_MIN_BLOCKS 2 with and without the word loads of the 4-pixel groups, and the encoder /
# golden / KAT suites on the word-load build.
mkdir -p gpurun_out
for rep in 1 2; do
for v in shipped mb2 mb2w mb3w; do
  if [ $v = shipped ]; then unset DVB_LIB_PATH; else export DVB_LIB_PATH=$PWD/_variants/libdvb_$v.so; fi
  timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c31_enc_${v}_$rep.json 2>/dev/null; echo "$v $rep: $(cut -c1-130 gpurun_out/c31_enc_${v}_$rep.json)"
done; done
export DVB_LIB_PATH=$PWD/_variants/libdvb_mb2w.so
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py tests/test_channel_planes.py tests/test_pair_support.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c31_pytest_mb2w.log 2>&1; echo "pytest mb2w exit $?"; tail -2 gpurun_out/c31_pytest_mb2w.log
